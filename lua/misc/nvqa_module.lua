-- Shared plumbing of the nn.Module shims lua/misc/LSTM.lua and lua/misc/netdef.lua.  WRITTEN BLIND (see nvqa_ffi.lua):
-- no Lua / Torch7 runtime exists in the build image; novel-vqa_b200/torch7_mirror.py is the executed twin of this file
-- (same classes, same entry points; tests/test_mirror_gpu.py::test_reference_shaped_jdj_through_module_calls).
--
-- Every shim module is an nn.Module over ONE block of a libnvqa model (block 0 = encoder_w_q, 1 = embedding_w_q,
-- 2 = multimodal_w, i.e. the three tensors of the reference's checkpoint):
--   :parameters()         -> {weight}, {gradWeight}   flat FloatTensors, so :getParameters() returns them unchanged
--   :updateOutput(input)  -> pushes weight (nvqa_params_set), stages the inputs, calls nvqa_*_forward
--   :backward(input, gradOutput) -> nvqa_grads_zero(block); nvqa_*_backward; gradWeight:add(nvqa_grads_get(block))
--   :zeroGradParameters(), :training(), :evaluate(), :clone(), :float(), :cuda() (no-ops: the math runs on the B200 either way)
-- Tensors are staged host -> device per call: this is the interface shim the reference's script sees; the throughput
-- path is nvqa_step.lua (one C call per training step).
local nvqa = require 'nvqa_ffi'
local ffi, lib = nvqa.ffi, nvqa.lib

local M = {}

-- one libnvqa model per (opt-derived) configuration, shared by the four modules of 002_train_baseline.lua:139-157
local shared = nil
function M.shared_model(cfg_table)
  if shared == nil then
    local cfg = ffi.new('nvqa_config', cfg_table)
    local h = ffi.new('nvqa_model*[1]')
    nvqa.check(lib.nvqa_model_create(cfg, h))
    shared = {h = ffi.gc(h[0], lib.nvqa_model_destroy), cfg = cfg}
  end
  return shared
end
function M.reset_shared_model() shared = nil end

-- device staging buffer holding a copy of a Float/IntTensor
local function dev_copy(model, t)
  local bytes = t:nElement() * t:elementSize()
  local p = ffi.new('void*[1]')
  nvqa.check(lib.nvqa_device_alloc(p, bytes))
  nvqa.check(lib.nvqa_memcpy_h2d(model.h, p[0], t:data(), bytes))
  return ffi.gc(p[0], lib.nvqa_device_free), bytes
end
local function dev_empty(bytes)
  local p = ffi.new('void*[1]')
  nvqa.check(lib.nvqa_device_alloc(p, bytes))
  return ffi.gc(p[0], lib.nvqa_device_free)
end
local function dev_fetch(model, p, t)
  nvqa.check(lib.nvqa_memcpy_d2h(model.h, t:data(), p, t:nElement() * t:elementSize()))
  return t
end
M.dev_copy, M.dev_empty, M.dev_fetch = dev_copy, dev_empty, dev_fetch

local Module, parent = torch.class('nvqa.Module', 'nn.Module')

function Module:__init(model, block)
  parent.__init(self)
  self.model, self.block = model, block
  local n = ffi.new('int64_t[1]')
  nvqa.check(lib.nvqa_param_count(model.h, block, n))
  self.weight = torch.FloatTensor(tonumber(n[0])):zero()
  self.gradWeight = torch.FloatTensor(tonumber(n[0])):zero()
  self.masks = nil          -- explicit Dropout multipliers (parity runs); nil = evaluate-mode call or hash masks
end

function Module:parameters() return {self.weight}, {self.gradWeight} end
function Module:zeroGradParameters() self.gradWeight:zero() end
function Module:push() nvqa.check(lib.nvqa_params_set(self.model.h, self.block, self.weight:data())) end
function Module:float() return self end
function Module:cuda() return self end

function Module:clone()                       -- dupe_rnn (misc/RNNUtils.lua:66-81): own (weight, gradWeight), same device model
  local c = {}
  for k, v in pairs(self) do c[k] = v end
  setmetatable(c, getmetatable(self))
  c.weight, c.gradWeight = self.weight:clone(), self.gradWeight:clone()
  return c
end

-- accGradParameters semantics: run one *_backward entry point on a zeroed gradient block, add the result to gradWeight
function Module:accumulate(call)
  nvqa.check(lib.nvqa_grads_zero(self.model.h, self.block))
  call()
  local g = torch.FloatTensor(self.gradWeight:size())
  nvqa.check(lib.nvqa_grads_get(self.model.h, self.block, g:data()))
  self.gradWeight:add(g)
end

function Module:mask_ptr(t)
  if not self.train or t == nil then return nil end
  local p = dev_copy(self.model, t:float():contiguous())
  return p
end

return M
