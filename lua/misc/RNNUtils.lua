-- Drop-in for 002_train_vqa_arch1/misc/RNNUtils.lua (functions on the arch1 hot path).  WRITTEN BLIND (see nvqa_ffi.lua).
-- right_align and sort_encoding_onehot_right_align keep the reference's names, argument meaning and 1-based
-- permutations; the dense one-hot [N x V] tensor is NOT materialised -- element [1] of the returned table is the
-- packed word-id vector (the one-hot nn.Linear is executed as a gather inside libnvqa).
local nvqa = require 'nvqa_ffi'
local ffi, lib = nvqa.ffi, nvqa.lib

function right_align(seq, lengths)            -- misc/RNNUtils.lua:54-61
  local s = seq:int():contiguous()
  local l = lengths:int():contiguous()
  local out = torch.IntTensor(s:size())
  nvqa.check(lib.nvqa_right_align(s:data(), l:data(), s:size(1), s:size(2), out:data()))
  return out:typeAs(seq)
end

function sort_encoding_onehot_right_align(batch_word_right_align, batch_length, vocabulary_size)   -- :84-125
  local q = batch_word_right_align:int():contiguous()
  local l = batch_length:int():contiguous()
  local B, T = q:size(1), q:size(2)
  local words, sizes = torch.IntTensor(B * T), torch.IntTensor(T)
  local sidx, inv = torch.IntTensor(B), torch.IntTensor(B)
  local nw, ns = ffi.new('int32_t[1]'), ffi.new('int32_t[1]')
  nvqa.check(lib.nvqa_pack_batch(q:data(), l:data(), B, T, words:data(), sizes:data(), sidx:data(), inv:data(), nw, ns))
  return {words:narrow(1, 1, nw[0]):long(), sizes:narrow(1, 1, ns[0]):long(), sidx:long(), inv:long()}
end

function join_vector(tensor_table)            -- :22-24
  return torch.cat(tensor_table, 1)
end

function split_vector(w, sizes)               -- :25-39
  local t, off = {}, 1
  local n = type(sizes) == 'table' and #sizes or sizes:size(1)
  for i = 1, n do
    t[#t + 1] = w[{{off, off + sizes[i] - 1}}]
    off = off + sizes[i]
  end
  return t
end

-- ---- the unrolled-RNN drivers (pure Lua over the module protocol; they run unchanged on the nvqa.LSTMCell shim) --------
function dupe_rnn(net, times)                 -- :66-81: {clones, their flat weights, their flat gradients}
  local nets, ws, dws = {}, {}, {}
  for i = 1, times do
    local c = net:clone()
    local w, dw = c:getParameters()
    nets[i], ws[i], dws[i] = c, w, dw
  end
  collectgarbage()
  return {nets, ws, dws}
end

-- :128-154, right-aligned batches (sizes non-decreasing): rows that become active start from init_state.  Unlike the
-- reference (App. C-2 of SURVEY.md) the grown state is a fresh tensor, not a view of init_state.
function rnn_forward(net_buffer, init_state, inputs, sizes)
  local N = sizes:size(1)
  local states = {init_state[{{1, sizes[1]}, {}}]}
  local outputs = {}
  for i = 1, N do
    if i > 1 and sizes[i] > sizes[i - 1] then
      local grown = init_state[{{1, sizes[i]}, {}}]:clone()
      grown[{{1, sizes[i - 1]}, {}}] = states[i]
      states[i] = grown
    elseif i > 1 and sizes[i] < sizes[i - 1] then
      error('left-aligned (shrinking) batches are not on the arch1 path')
    end
    states[i + 1] = net_buffer[1][i]:forward({states[i], inputs[i]})
  end
  return states, outputs
end

-- :181-210, the branch JdJ takes (doutputs is the dummy output tensor, not a table)
function rnn_backward(net_buffer, dend_state, doutputs, states, inputs, sizes)
  local N = sizes:size(1)
  local dstate = {[N + 1] = dend_state[{{1, sizes[N]}, {}}]}
  local dinputs = {}
  for i = N, 1, -1 do
    local g = net_buffer[1][i]:backward({states[i], inputs[i]}, dstate[i + 1])
    dstate[i] = (i == 1 or sizes[i] == sizes[i - 1]) and g[1] or g[1][{{1, sizes[i - 1]}, {}}]
    dinputs[i] = g[2]
  end
  return dstate, dinputs
end
