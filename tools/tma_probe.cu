// tma_probe.cu — development probe: one warp, one TMA 2-D tile load of a uint8 window.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tma_probe.cu && ./tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int variant, int x, int y,
                        int box_w, int box_h, unsigned char *out, int *status) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem + 4096);
  const uint32_t bar_u = smem_u32(bar), dst = smem_u32(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_u), "r"(1) : "memory");
    if (variant != 4) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (variant == 1) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_u) : "memory");
    } else {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u), "r"(box_w * box_h) : "memory");
      const CUtensorMap *tp = variant == 2 ? gmap : &tmap;
      if (variant == 3)
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(dst), "l"(tp), "r"(x), "r"(y), "r"(bar_u) : "memory");
      else
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
            ::"r"(dst), "l"(tp), "r"(x), "r"(y), "r"(bar_u) : "memory");
    }
  }
  int ok = 0;
  for (uint32_t spin = 0; spin < (1u << 20) && !ok; ++spin) {
    uint32_t r;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(r) : "r"(bar_u), "r"(0) : "memory");
    ok = r;
  }
  if (threadIdx.x == 0) *status = ok;
  __syncwarp();
  for (int i = threadIdx.x; i < box_w * box_h; i += 32) out[i] = smem[i];
}

int main(int argc, char **argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 256, H = 192, box_w = 32, box_h = 28;
  std::vector<unsigned char> h(W * H);
  for (int i = 0; i < W * H; ++i) h[i] = (unsigned char)((i * 7 + i / W) & 0xff);
  unsigned char *d_map, *d_out;
  int *d_status;
  cudaMalloc(&d_map, W * H);
  cudaMalloc(&d_out, box_w * box_h);
  cudaMalloc(&d_status, 4);
  cudaMemcpy(d_map, h.data(), W * H, cudaMemcpyHostToDevice);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fn);
  typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {W, H}, gstride[1] = {W};
  const cuuint32_t box[2] = {box_w, box_h}, es[2] = {1, 1};
  CUresult r = ((Enc)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_map, gdim, gstride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  CUtensorMap *d_tm;
  cudaMalloc(&d_tm, sizeof(CUtensorMap));
  cudaMemcpy(d_tm, &tm, sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  printf("variant %d\n", variant);
  for (int t = 0; t < 3; ++t) {
    const int xs[3] = {10, 240, 5}, ys[3] = {20, 180, 0};
    k_probe<<<1, 32, 8192>>>(tm, d_tm, variant, xs[t], ys[t], box_w, box_h, d_out, d_status);
    e = cudaDeviceSynchronize();
    int st = -1;
    std::vector<unsigned char> o(box_w * box_h);
    cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o.data(), d_out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int yy = 0; yy < box_h; ++yy)
      for (int xx = 0; xx < box_w; ++xx) {
        const int gx = xs[t] + xx, gy = ys[t] + yy;
        const unsigned char want = (gx < W && gy < H) ? h[gy * W + gx] : 0;
        bad += o[yy * box_w + xx] != want;
      }
    printf("probe (%d,%d): sync=%s status=%d mismatches=%d\n", xs[t], ys[t], cudaGetErrorString(e), st, bad);
  }
  return 0;
}
