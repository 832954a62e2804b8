-- The hot path of 002_train_vqa_arch1/002_train_baseline.lua behind the reference's own call shape.
-- WRITTEN BLIND (see nvqa_ffi.lua).  Usage inside the training script, replacing :141-181 (model construction),
-- :272-335 (JdJ) and :408 (optim.rmsprop):
--
--   local step = require 'nvqa_step'
--   local model = step.create(opt, vocabulary_size_q, buffer_size_q)        -- nets + getParameters()
--   model:set_params{encoder_w_q = ..., embedding_w_q = ..., multimodal_w = ...}   -- uniform(-0.08, 0.08) or torch.load
--   for iter = 1, opt.max_iters do
--     local q, len, fv_im, labels = next_batch_raw()       -- dataset:next_batch() without the sort / one-hot
--     local f = model:train_step(q, len, fv_im, labels, optimize.learningRate, opt.seed + iter)
--     running_avg = running_avg and (running_avg * 0.95 + f * 0.05) or f     -- :330-333
--     optimize.learningRate = optimize.learningRate * decay_factor            -- :410
--   end
--   torch.save(path, model:get_params())                   -- same 3-tensor table as :401-402
local nvqa = require 'nvqa_ffi'
local ffi, lib = nvqa.ffi, nvqa.lib
local BLOCKS = {encoder_w_q = 0, embedding_w_q = 1, multimodal_w = 2}

local Model = {}
Model.__index = Model

local function create(opt, vocabulary_size, T)
  local cfg = ffi.new('nvqa_config', {arch = 1, V = vocabulary_size, E = opt.input_encoding_size, H = opt.rnn_size,
    L = opt.rnn_layer, I = opt.nhimage, C = opt.common_embedding_size, O = opt.num_output, T = T, B = opt.batch_size,
    precision = 3 --[[NVQA_PREC_BF16X2]], img_norm = 0 --[[the script normalises at load time, :117-123]],
    device = opt.gpuid, dropout = 0.5})
  local h = ffi.new('nvqa_model*[1]')
  nvqa.check(lib.nvqa_model_create(cfg, h))
  return setmetatable({h = ffi.gc(h[0], lib.nvqa_model_destroy), cfg = cfg}, Model)
end

function Model:set_params(t)
  for name, blk in pairs(BLOCKS) do
    local w = t[name]:float():contiguous()
    nvqa.check(lib.nvqa_params_set(self.h, blk, w:data()))
  end
end

function Model:get_params()
  local out, n = {}, ffi.new('int64_t[1]')
  for name, blk in pairs(BLOCKS) do
    nvqa.check(lib.nvqa_param_count(self.h, blk, n))
    local w = torch.FloatTensor(tonumber(n[0]))
    nvqa.check(lib.nvqa_params_get(self.h, blk, w:data()))
    out[name] = w
  end
  return out
end

-- q: IntTensor [B x T] right-aligned, len: IntTensor [B], fv_im: FloatTensor [B x I], labels: IntTensor [B]
function Model:train_step(q, len, fv_im, labels, lr, seed)
  local loss = ffi.new('float[1]')
  nvqa.check(lib.nvqa_train_step_host(self.h, q:data(), len:data(), fv_im:data(), labels:data(), q:size(1), lr, seed, loss))
  return loss[0]
end

-- 004_eval_model.lua:202-233: forward + torch.max(scores, 2)
function Model:predict(q, len, fv_im)
  local ans = torch.IntTensor(q:size(1))
  nvqa.check(lib.nvqa_eval_step_host(self.h, q:data(), len:data(), fv_im:data(), q:size(1), ans:data()))
  return ans:long()
end

-- ---------------------------------------------------------------------------------------------------------------
-- Text autoencoder (001_train_autoencoder/001_train_arch1_text_autoencoder.lua): replaces protos.ae / protos.crit
-- construction (:95-108), lossFun (:208-249) and the adam call (:334).
--   local ae = step.create_ae(opt, loader:getVocabSize(), loader:getSeqLength())
--   ae:set_params{encoder = ..., decoder = ..., lookup_table = ...}
--   local loss = ae:train_step(data.labels:t():int():contiguous(), lengths, learning_rate, opt.seed + iter)
local AE_BLOCKS = {encoder = 0, decoder = 1, lookup_table = 2}
local AE = setmetatable({}, {__index = Model})
AE.__index = AE

local function create_ae(opt, vocab_size, seq_length)
  local cfg = ffi.new('nvqa_config', {arch = 3, V = vocab_size, E = opt.input_encoding_size, H = opt.rnn_size,
    L = opt.num_layers, I = 4, C = 0, O = 4, T = seq_length, B = opt.batch_size, precision = 3, img_norm = 0,
    device = opt.gpuid, dropout = opt.drop_prob_ae})
  local h = ffi.new('nvqa_model*[1]')
  nvqa.check(lib.nvqa_model_create(cfg, h))
  return setmetatable({h = ffi.gc(h[0], lib.nvqa_model_destroy), cfg = cfg}, AE)
end

function AE:set_params(t)
  for name, blk in pairs(AE_BLOCKS) do
    nvqa.check(lib.nvqa_params_set(self.h, blk, t[name]:float():contiguous():data()))
  end
end

-- seq: IntTensor [B x T] (data.labels transposed), zero-padded on the right; len: IntTensor [B]
function AE:train_step(seq, len, lr, seed)
  local loss = ffi.new('float[1]')
  nvqa.check(lib.nvqa_train_step_host(self.h, seq:data(), len:data(), nil, nil, seq:size(1), lr, seed, loss))
  return loss[0]
end

return {create = create, create_ae = create_ae}
