-- Drop-in for 002_train_vqa_arch1/misc/netdef.lua.  WRITTEN BLIND (see nvqa_ffi.lua); executed twin:
-- novel-vqa_b200/torch7_mirror.py::_AxB / _MultimodalNet / _EmbeddingNet.
--
-- netdef.AxB(nhA, nhB, nhcommon, dropout) / netdef.AskipB(...) return an nn.Module with the gModule's protocol
-- (misc/netdef.lua:6-25):
--   forward({q, i})            -> tanh(Wq drop(q)) (.) tanh(Wi drop(i))            [rows x nhcommon]
--   backward({q, i}, dout)     -> {dq, di};  gradWeight += d(Wq, bq, Wi, bi)
-- netdef.multimodal(model) is the whole multimodal_net of 002_train_baseline.lua:151-154 (AxB + Dropout + Linear(C, O)) in
-- one module (one C call per forward / backward instead of three modules), netdef.embedding(model) the embedding_net_q of
-- :141-144 on the packed word ids that sort_encoding_onehot_right_align (lua/misc/RNNUtils.lua) returns in place of the
-- dense one-hot matrix.  All of them own block 2 / block 1 of the shared libnvqa model.
local nvqa = require 'nvqa_ffi'
local base = require 'misc.nvqa_module'
local ffi, lib = nvqa.ffi, nvqa.lib

netdef = {}

-- ---- netdef.AxB / AskipB ------------------------------------------------------------------------------------------
local AxB, parent = torch.class('nvqa.AxB', 'nvqa.Module')

function AxB:__init(model, skip)
  parent.__init(self, model, 2 --[[NVQA_BLOCK_MULTIMODAL]])
  nvqa.check(lib.nvqa_set_variant(model.h, skip and 1 or 0, 1.0, 0))
end

-- masks = {mask_q [rows x nhA], mask_i [rows x nhB]} (+ mask_z [rows x nhcommon] for the multimodal module)
local function three_masks(self)
  local m = self.masks
  if not self.train then return nil, nil, nil end
  if m == nil then error('training-mode forward needs explicit .masks') end
  return self:mask_ptr(m[1]), self:mask_ptr(m[2]), m[3] and self:mask_ptr(m[3]) or nil
end

function AxB:updateOutput(input)
  local q, i = input[1]:float():contiguous(), input[2]:float():contiguous()
  self:push()
  local rows = q:size(1)
  local Q_, I_ = base.dev_copy(self.model, q), base.dev_copy(self.model, i)
  local O_ = base.dev_empty(rows * self.model.cfg.C * 4)
  local mq, mi = three_masks(self)
  nvqa.check(lib.nvqa_axb_forward(self.model.h, Q_, I_, mq, mi, rows, O_))
  self.output = base.dev_fetch(self.model, O_, torch.FloatTensor(rows, self.model.cfg.C))
  return self.output
end

function AxB:backward(input, gradOutput)
  local q, i = input[1]:float():contiguous(), input[2]:float():contiguous()
  local g = gradOutput:float():contiguous()
  self:push()
  local rows = q:size(1)
  local Q_, I_, G_ = base.dev_copy(self.model, q), base.dev_copy(self.model, i), base.dev_copy(self.model, g)
  local DQ_, DI_ = base.dev_empty(q:nElement() * 4), base.dev_empty(i:nElement() * 4)
  local mq, mi = three_masks(self)
  self:accumulate(function()
    nvqa.check(lib.nvqa_axb_backward(self.model.h, Q_, I_, mq, mi, G_, rows, DQ_, DI_))
  end)
  self.gradInput = {base.dev_fetch(self.model, DQ_, torch.FloatTensor(q:size())),
                    base.dev_fetch(self.model, DI_, torch.FloatTensor(i:size()))}
  return self.gradInput
end
AxB.updateGradInput = AxB.backward
function AxB:accGradParameters() end

-- ---- multimodal_net = Sequential{AxB, Dropout, Linear(C, O)} ------------------------------------------------------
local MM, mmparent = torch.class('nvqa.Multimodal', 'nvqa.Module')

function MM:__init(model)
  mmparent.__init(self, model, 2)
end

function MM:updateOutput(input)
  local q, i = input[1]:float():contiguous(), input[2]:float():contiguous()
  self:push()
  local rows = q:size(1)
  local Q_, I_ = base.dev_copy(self.model, q), base.dev_copy(self.model, i)
  local O_ = base.dev_empty(rows * self.model.cfg.O * 4)
  local mq, mi, mz = three_masks(self)
  nvqa.check(lib.nvqa_multimodal_forward(self.model.h, Q_, I_, mq, mi, mz, rows, O_))
  self.output = base.dev_fetch(self.model, O_, torch.FloatTensor(rows, self.model.cfg.O))
  return self.output
end

function MM:backward(input, gradOutput)
  local q, i = input[1]:float():contiguous(), input[2]:float():contiguous()
  local g = gradOutput:float():contiguous()
  self:push()
  local rows = q:size(1)
  local Q_, I_, G_ = base.dev_copy(self.model, q), base.dev_copy(self.model, i), base.dev_copy(self.model, g)
  local DQ_, DI_ = base.dev_empty(q:nElement() * 4), base.dev_empty(i:nElement() * 4)
  local mq, mi, mz = three_masks(self)
  self:accumulate(function()
    nvqa.check(lib.nvqa_multimodal_backward(self.model.h, Q_, I_, mq, mi, mz, G_, rows, DQ_, DI_))
  end)
  self.gradInput = {base.dev_fetch(self.model, DQ_, torch.FloatTensor(q:size())),
                    base.dev_fetch(self.model, DI_, torch.FloatTensor(i:size()))}
  return self.gradInput
end
MM.updateGradInput = MM.backward
function MM:accGradParameters() end

-- ---- embedding_net_q = Sequential{Linear(V, E), Dropout, Tanh} on one-hot rows -------------------------------------
local Emb, embparent = torch.class('nvqa.Embedding', 'nvqa.Module')

function Emb:__init(model)
  embparent.__init(self, model, 1 --[[NVQA_BLOCK_EMBEDDING]])
end

-- words: LongTensor [N], element [1] of sort_encoding_onehot_right_align; masks: FloatTensor [N x E]
function Emb:updateOutput(words)
  local w = words:int():contiguous()
  self:push()
  local n = w:nElement()
  local W_ = base.dev_copy(self.model, w)
  local Y_ = base.dev_empty(n * self.model.cfg.E * 4)
  nvqa.check(lib.nvqa_embedding_forward(self.model.h, W_, self:mask_ptr(self.masks), n, Y_))
  self.output = base.dev_fetch(self.model, Y_, torch.FloatTensor(n, self.model.cfg.E))
  return self.output
end

function Emb:backward(words, gradOutput)
  local w = words:int():contiguous()
  local g = gradOutput:float():contiguous()
  self:push()
  local n = w:nElement()
  local W_, Y_, G_ = base.dev_copy(self.model, w), base.dev_copy(self.model, self.output), base.dev_copy(self.model, g)
  local mk = self:mask_ptr(self.masks)
  self:accumulate(function()
    nvqa.check(lib.nvqa_embedding_backward(self.model.h, W_, Y_, G_, mk, n))
  end)
  self.gradInput = nil                          -- the [N x V] gradInput is never read by the reference (:320)
  return self.gradInput
end
Emb.updateGradInput = Emb.backward
function Emb:accGradParameters() end

-- ---- constructors with the reference's names ----------------------------------------------------------------------
local function model_for(nhA, nhB, nhcommon, dropout)
  local cfg = {}
  for k, v in pairs(LSTM and LSTM.config or {arch = 1, V = 8, E = 4, T = 26, B = 500, O = 4, precision = 3, img_norm = 0, device = 0}) do cfg[k] = v end
  cfg.I, cfg.C, cfg.dropout = nhB, nhcommon, dropout
  return base.shared_model(cfg)               -- rnn_size / rnn_layer come from LSTM.config: nhA = 2 * H * L
end

function netdef.AxB(nhA, nhB, nhcommon, dropout)            -- misc/netdef.lua:6-14
  return nvqa.AxB(model_for(nhA, nhB, nhcommon, dropout or 0), false)
end
function netdef.AskipB(nhA, nhB, nhcommon, dropout)         -- misc/netdef.lua:16-25
  return nvqa.AxB(model_for(nhA, nhB, nhcommon, dropout or 0), true)
end
function netdef.multimodal(model) return nvqa.Multimodal(model) end      -- 002_train_baseline.lua:151-154
function netdef.embedding(model) return nvqa.Embedding(model) end        -- 002_train_baseline.lua:141-144

return netdef
