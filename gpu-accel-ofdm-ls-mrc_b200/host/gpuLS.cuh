// gpuLS.cuh -- forwarding header: the reference's GPU function surface under its own file name (gpuLS.cuh:72-113 and
// the free-function spellings gpuLS_main.cu:104-141 calls), implemented by gpuLS.hpp on top of the C ABI.  With this
// and ShMemSymBuff_cucomplex.hpp on the include path, /root/reference/gpuLS_main.cu builds unmodified
// (oracle/Makefile, target _ref/gpuLS_main_ref_%; tests/test_gpu_golden_and_host.py runs it against the oracle).
#ifndef LSMRC_HOST_GPULS_CUH_
#define LSMRC_HOST_GPULS_CUH_

#include <cuComplex.h>
#include <cuda_runtime.h>

#include "gpuLS.hpp"

#endif
