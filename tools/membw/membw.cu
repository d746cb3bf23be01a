// Read-bandwidth probe (tools only, not part of the product library): how do access pattern, CTA size and
// loads in flight per thread change the achieved HBM read rate on B200?
#include <cuda_runtime.h>
#include <cstdint>

template <int U>
__global__ void read_stream(const uint4* __restrict__ p, size_t n16, unsigned* sink) {
  // grid-stride: at any instant the whole grid touches one contiguous window
  unsigned acc = 0;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + (U - 1) * stride < n16; i += U * stride) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = p[i + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

template <int U>
__global__ void read_region(const uint4* __restrict__ p, size_t n16_per_cta, unsigned* sink) {
  // every CTA streams its own contiguous region (the shape of the per-sample kernels)
  unsigned acc = 0;
  const uint4* q = p + blockIdx.x * n16_per_cta;
  size_t i = threadIdx.x;
  for (; i + (U - 1) * (size_t)blockDim.x < n16_per_cta; i += U * (size_t)blockDim.x) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = q[i + u * blockDim.x];
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

extern "C" int membw_launch(int mode, int unroll, const void* p, size_t bytes, int grid, int block, void* sink, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n16 = bytes / 16;
  const uint4* q = (const uint4*)p;
  unsigned* k = (unsigned*)sink;
  if (mode == 0) {
    if (unroll == 1) read_stream<1><<<grid, block, 0, s>>>(q, n16, k);
    else if (unroll == 2) read_stream<2><<<grid, block, 0, s>>>(q, n16, k);
    else if (unroll == 4) read_stream<4><<<grid, block, 0, s>>>(q, n16, k);
    else read_stream<8><<<grid, block, 0, s>>>(q, n16, k);
  } else {
    const size_t per = n16 / grid;
    if (unroll == 1) read_region<1><<<grid, block, 0, s>>>(q, per, k);
    else if (unroll == 2) read_region<2><<<grid, block, 0, s>>>(q, per, k);
    else if (unroll == 4) read_region<4><<<grid, block, 0, s>>>(q, per, k);
    else read_region<8><<<grid, block, 0, s>>>(q, per, k);
  }
  return (int)cudaGetLastError();
}
