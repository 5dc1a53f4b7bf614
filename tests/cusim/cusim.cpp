// tests/cusim/cusim.cpp -- storage for the host-emulation thread context (test infrastructure only, see cusim.h)
#include "cusim.h"
namespace cusim {
thread_local uint3_ t_threadIdx, t_blockIdx;
thread_local dim3 t_blockDim, t_gridDim;
thread_local BlockCtx *t_ctx;
thread_local WarpCtx *t_warp;
thread_local unsigned t_lane, t_xpar;
}
