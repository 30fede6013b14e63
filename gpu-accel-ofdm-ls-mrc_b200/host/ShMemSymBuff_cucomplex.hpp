// ShMemSymBuff_cucomplex.hpp -- forwarding header: lets a caller written against the reference's CUDA-flavoured ring
// header (`#include "ShMemSymBuff_cucomplex.hpp"`, gpuLS_main.cu:34) compile unchanged against this repo.
// The reference header includes <cuComplex.h> itself (ShMemSymBuff_cucomplex.hpp:36) and DEFINES two globals that
// its driver uses without declaring them: the output stream `outfile` (:73, gpuLS_main.cu:120-125) and the repeat
// count `numTimes` (:87, gpuLS_main.cu:106).  Both are reproduced here; everything else comes from ShMemSymBuff.hpp.
#ifndef LSMRC_HOST_SHMEMSYMBUFF_CUCOMPLEX_HPP_
#define LSMRC_HOST_SHMEMSYMBUFF_CUCOMPLEX_HPP_

#include <cuComplex.h>
#include <cuda_runtime.h>

#include <fstream>

#ifndef cudaEn
#define cudaEn
#endif
#include "ShMemSymBuff.hpp"

std::ofstream outfile;  // (one translation unit per program, as in the reference)
float numTimes = 1;

#endif
