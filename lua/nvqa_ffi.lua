-- LuaJIT FFI binding of libnvqa.so (include/nvqa.h).  WRITTEN BLIND: no Lua / LuaJIT / Torch7 runtime exists in the
-- build image (SURVEY F3), so this file has never been executed; the same declarations are exercised through
-- Python ctypes (novel-vqa_b200/_lib.py, tests/test_abi.py).
local ffi = require 'ffi'

-- The header is restricted to the C subset ffi.cdef accepts; only the include guard / extern "C" lines are dropped.
local function read_header(path)
  local f = assert(io.open(path, 'r'))
  local out = {}
  for line in f:lines() do
    if not line:match('^%s*#') and not line:match('extern "C"') and not line:match('^}%s*$') then
      out[#out + 1] = line
    end
  end
  f:close()
  return table.concat(out, '\n')
end

local root = os.getenv('NVQA_ROOT') or '.'
ffi.cdef('typedef signed char int8_t; typedef int int32_t; typedef long long int64_t; typedef unsigned long long uint64_t;')
ffi.cdef(read_header(root .. '/include/nvqa.h'))
local lib = ffi.load(root .. '/novel-vqa_b200/csrc/libnvqa.so')

local M = {lib = lib, ffi = ffi}

-- Torch7 convention: errors are raised with error() (THError); scripts never pcall.
function M.check(rc)
  if rc ~= 0 then error('libnvqa: ' .. ffi.string(lib.nvqa_last_error()), 2) end
end

return M
