// Device-resident ensembles: the realizations of `rand(process, domain, nreals)` stay in HBM,
// sharded contiguously over the devices of the context, and the per-node statistics of src/ensembles.jl:42-52
// run over them in place.  Shared between ensemble.cu (statistics) and fft.cu / lusim.cu (filling by simulation).
#pragma once
#include <memory>

#include "common.h"

namespace gsp {

struct EnsDev {
  DevCtx* dc = nullptr;
  DevBuf Z;              // n x nr doubles, realization index slowest (column r of the reference's table)
  long long r0 = 0, nr = 0;  // realizations [r0, r0 + nr) of the ensemble live here
};

}  // namespace gsp

struct gsp_ensemble {
  gsp_ctx* ctx = nullptr;
  long long n = 0, R = 0;
  std::vector<std::unique_ptr<gsp::EnsDev>> dev;
  std::mutex mu;
};
