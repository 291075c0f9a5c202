// CUDA programming-guide TMA sample (libcu++ wrappers), int32 tensor, to check that TMA works at all here.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;

template <typename T, int BW, int BH>
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, T *out) {
  __shared__ alignas(128) T smem_buffer[BH][BW];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) {
    init(&bar, blockDim.x);
    cde::fence_proxy_async_shared_cta();
  }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
  } else {
    token = bar.arrive();
  }
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = smem_buffer[i / BW][i % BW];
}

template <typename T, int BW, int BH>
int run(CUtensorMapDataType dt, int W, int H) {
  std::vector<T> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (T)(i * 7 + i / W);
  T *d_map, *d_out;
  cudaMalloc(&d_map, sizeof(T) * W * H);
  cudaMalloc(&d_out, sizeof(T) * BW * BH);
  cudaMemcpy(d_map, h.data(), sizeof(T) * W * H, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}, gstride[1] = {(cuuint64_t)W * sizeof(T)};
  const cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&tm, dt, 2, d_map, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int X = getenv("PX") ? atoi(getenv("PX")) : 8;
  kernel<T, BW, BH><<<1, 128>>>(tm, X, 4, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<T> o(BW * BH);
  cudaMemcpy(o.data(), d_out, sizeof(T) * o.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int yy = 0; yy < BH; ++yy)
    for (int xx = 0; xx < BW; ++xx) bad += o[yy * BW + xx] != h[(size_t)(4 + yy) * W + X + xx];
  printf("T=%zuB box %dx%d: encode=%d sync=%s mismatches=%d\n", sizeof(T), BW, BH, (int)r, cudaGetErrorString(e), bad);
  return 0;
}

int main(int argc, char **argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 0;
  if (v == 0) return run<int, 64, 64>(CU_TENSOR_MAP_DATA_TYPE_INT32, 256, 256);
  if (v == 1) return run<unsigned char, 64, 64>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 256, 256);
  if (v == 2) return run<unsigned char, 32, 28>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 256, 192);
  if (v == 3) return run<unsigned char, 32, 32>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 256, 192);
  if (v == 4) return run<unsigned char, 128, 28>(CU_TENSOR_MAP_DATA_TYPE_UINT8, 256, 192);
  return 0;
}
