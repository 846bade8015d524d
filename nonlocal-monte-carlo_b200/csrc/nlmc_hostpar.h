// nlmc_hostpar.h -- host-side worker pool and format conversions (nlmc_hostpar.cpp); no CUDA types.
#pragma once
#include <cstdint>
#include <functional>

namespace nlmc {

int host_threads();
int host_threads_shared();  // per-rank share of the pool in a multi-process job (LOCAL_WORLD_SIZE)
// fn(part, parts) on min(parts, host_threads()) persistent workers; returns when every part is done
void parallel_for(int parts, const std::function<void(int, int)> &fn);
// out[i] = (double)in[i] on `threads` workers (0 = all), non-temporal stores where the CPU has AVX2
void widen_i8_f64(const int8_t *in, double *out, uint64_t count, int threads);
// first-touch every page of a fresh buffer on `threads` workers (0 = all)
void prefault(void *buf, uint64_t bytes, int threads);

}  // namespace nlmc
