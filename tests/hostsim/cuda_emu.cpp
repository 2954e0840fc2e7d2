// storage for the emulator's builtin variables (TEST INFRASTRUCTURE ONLY, see cuda_emu.h)
#include "cuda_emu.h"
namespace emu {
thread_local uint3 g_threadIdx = {0, 0, 0}, g_blockIdx = {0, 0, 0};
thread_local dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);
}  // namespace emu
