-- Drop-in for 002_train_vqa_arch1/misc/LSTM.lua.  WRITTEN BLIND (see nvqa_ffi.lua); executed twin:
-- novel-vqa_b200/torch7_mirror.py::_LSTMCell.
--
-- LSTM.lstm_conventional(input_size, rnn_size, noutput, n, dropout) returns -- instead of the nngraph gModule of
-- misc/LSTM.lua:12-73 -- an nn.Module with the same call protocol:
--   forward({state, x})            -> state'      state = [c1 h1 c2 h2 ...] [rows x 2*n*rnn_size]   (:15-72)
--   backward({state, x}, dstate')  -> {dstate, dx}; gradWeight += this clone's dW   (misc/RNNUtils.lua:195-196)
--   getParameters() -> flat w, dw in nngraph order: per layer i2h.weight, i2h.bias, h2h.weight, h2h.bias (:41-42)
--   clone(), training(), evaluate(), zeroGradParameters()
-- so dupe_rnn / rnn_forward / rnn_backward of the reference's own misc/RNNUtils.lua run unchanged on top of it.
-- The gate order (i, f, o, g rows of the 4*rnn_size Linear outputs, :45-52) and the inter-layer Dropout (:36-37) are the
-- library's; `noutput` is unused by the reference too (:66-68).
local nvqa = require 'nvqa_ffi'
local base = require 'misc.nvqa_module'
local ffi, lib = nvqa.ffi, nvqa.lib

LSTM = {}

local Cell, parent = torch.class('nvqa.LSTMCell', 'nvqa.Module')

function Cell:__init(model, input_size, rnn_size, n, dropout)
  parent.__init(self, model, 0 --[[NVQA_BLOCK_ENCODER]])
  self.input_size, self.rnn_size, self.n, self.dropout = input_size, rnn_size, n, dropout
end

local function check_sizes(self, state, x)
  if state:size(2) ~= 2 * self.n * self.rnn_size or x:size(2) ~= self.input_size or state:size(1) ~= x:size(1) then
    error('size mismatch')                      -- what THNN's Linear raises
  end
end

-- masks: FloatTensor [(n-1) x rows x rnn_size] of Dropout multipliers (0 or 1/(1-p)); required for training-mode calls
-- with n > 1 because Torch7's RNG stream cannot be reproduced inside the library one module call at a time
function Cell:updateOutput(input)
  local state, x = input[1]:float():contiguous(), input[2]:float():contiguous()
  check_sizes(self, state, x)
  self:push()
  local rows = state:size(1)
  local S_, X_ = base.dev_copy(self.model, state), base.dev_copy(self.model, x)
  local O_ = base.dev_empty(state:nElement() * 4)
  local mk = (self.n > 1 and self.dropout > 0) and self:mask_ptr(self.masks) or nil
  if self.train and self.n > 1 and self.dropout > 0 and mk == nil then error('training-mode forward needs explicit .masks') end
  nvqa.check(lib.nvqa_lstm_cell_forward(self.model.h, S_, X_, mk, rows, O_))
  self.output = base.dev_fetch(self.model, O_, torch.FloatTensor(state:size()))
  return self.output
end

function Cell:backward(input, gradOutput)
  local state, x = input[1]:float():contiguous(), input[2]:float():contiguous()
  local g = gradOutput:float():contiguous()
  check_sizes(self, state, x)
  self:push()
  local rows = state:size(1)
  local S_, X_, G_ = base.dev_copy(self.model, state), base.dev_copy(self.model, x), base.dev_copy(self.model, g)
  local DS_, DX_ = base.dev_empty(state:nElement() * 4), base.dev_empty(x:nElement() * 4)
  local mk = (self.n > 1 and self.dropout > 0) and self:mask_ptr(self.masks) or nil
  self:accumulate(function()
    nvqa.check(lib.nvqa_lstm_cell_backward(self.model.h, S_, X_, mk, G_, rows, DS_, DX_))
  end)
  self.gradInput = {base.dev_fetch(self.model, DS_, torch.FloatTensor(state:size())),
                    base.dev_fetch(self.model, DX_, torch.FloatTensor(x:size()))}
  return self.gradInput
end
Cell.updateGradInput = Cell.backward          -- nn.Module:backward = updateGradInput + accGradParameters, fused here
function Cell:accGradParameters() end

-- The libnvqa model behind the cell is the one shared by all arch1 modules (misc/nvqa_module.lua); the trainer fills
-- LSTM.config (V, I, C, O, T, B ...) from `opt` before constructing the nets, see INTEGRATION.md.
LSTM.config = {arch = 1, V = 8, I = 4, C = 4, O = 4, T = 26, B = 500, precision = 3, img_norm = 0, device = 0, dropout = 0.5}

function LSTM.lstm_conventional(input_size, rnn_size, noutput, n, dropout)
  dropout = dropout or 0
  local cfg = {}
  for k, v in pairs(LSTM.config) do cfg[k] = v end
  cfg.E, cfg.H, cfg.L, cfg.dropout = input_size, rnn_size, n, dropout
  return nvqa.LSTMCell(base.shared_model(cfg), input_size, rnn_size, n, dropout)
end

return LSTM
