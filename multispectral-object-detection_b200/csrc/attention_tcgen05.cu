// 128-token multi-head self-attention core on the tensor cores (sm_100a), one CTA per (image, head):
//
//   S = Q K^T      tcgen05.mma  128 x 128 x dk   (A = Q, B = K, both K-major smem tiles loaded by TMA
//                                                 straight out of the fused QKV activation [B*128, 3C])
//   P = softmax(S / sqrt(dk))   128 threads, thread t = TMEM lane t = query row t; fp32; exp2f with the
//                               scale folded in; P written to smem as the bf16 K-major A operand
//   O = P V        tcgen05.mma  128 x dk x 128   (B = V as loaded: [key][dk] = MN-major operand)
//   out = O / rowsum  -> bf16, heads merged ([B*128, C])
//
// Swizzle mode of the Q/K/V tiles follows the head dim: dk*2 bytes per row -> 32B / 64B / 128B
// (dk = 128: two 64-wide chunks).  Replaces models/common.py:497-510 (two batched matmuls + softmax that
// materialise the [B,8,128,128] attention tensor in HBM).
#include "cft_common.cuh"
#include "tcgen05_ptx.cuh"

namespace {
using namespace cft;
using namespace cft::ptx;

constexpr int kT = 128;          // tokens
constexpr int kThreads = 128;

struct AttnParams {
  int C, heads, dk;
  int bw;          // box width in elements = min(dk, 64)
  int nchunk;      // dk / bw
  int layout;      // UMMA layout type of the Q/K/V tiles (2 = SW128, 4 = SW64, 6 = SW32)
  float scale_log2e;
  __nv_bfloat16* out;
};

__global__ void __launch_bounds__(kThreads)
cft_attention_tcgen05_kernel(const __grid_constant__ CUtensorMap qkv_map, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row_bytes = p.bw * 2;
  const int chunk_bytes = kT * row_bytes;            // one [128 x bw] tile
  const int tile_bytes = chunk_bytes * p.nchunk;     // one of Q / K / V
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + tile_bytes;
  uint8_t* sV = sK + tile_bytes;
  // P: 2 chunks x [128 x 64] bf16, SWIZZLE_128B (32 KiB).  It is written only after S = Q K^T has retired, so for
  // dk >= 64 it reuses the Q|K tiles (dead by then): 97 KiB per CTA at dk = 128 -> two CTAs per SM, and the 256
  // (image, head) CTAs of a batch-32 launch run as one wave instead of two.
  const bool alias_p = 2 * tile_bytes >= 2 * kT * 128;
  uint8_t* sP = alias_p ? sQ : sV + tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + tile_bytes + (alias_p ? 0 : 2 * kT * 128));
  uint64_t* tma_bar = bars;
  uint64_t* mma_bar = bars + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  if (tid == 0) {
    prefetch_tmap(&qkv_map);
    mbar_init(tma_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;          // S: columns [0,128)
  const uint32_t tmem_o = tmem_base + 128;    // O: columns [128,128+dk)
  // barrier init / TMEM allocation above overlap the tail of the QKV GEMM; its output is read from here on
  pdl_launch_dependents();
  pdl_wait();

  if (tid == 0) {
    mbar_arrive_expect_tx(tma_bar, 3u * static_cast<uint32_t>(tile_bytes));
    const int r0 = b * kT;
    for (int c = 0; c < p.nchunk; ++c) {
      const int col = h * p.dk + c * p.bw;
      tma_load_2d(sQ + c * chunk_bytes, &qkv_map, tma_bar, col, r0);
      tma_load_2d(sK + c * chunk_bytes, &qkv_map, tma_bar, p.C + col, r0);
      tma_load_2d(sV + c * chunk_bytes, &qkv_map, tma_bar, 2 * p.C + col, r0);
    }
  }
  mbar_wait(tma_bar, 0);
  tc_fence_after();

  // ---- S = Q K^T : K-major operands; 8-row groups are 8*row_bytes apart (SBO) ----
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_ex(128, 128, 0, 0);
    const uint32_t sbo = 8u * row_bytes;
    int kk = 0;
    for (int c = 0; c < p.nchunk; ++c) {
      for (int k = 0; k < p.bw / 16; ++k, ++kk) {
        const uint64_t da = umma_desc(smem_u32(sQ + c * chunk_bytes) + k * 32, 0, sbo, p.layout);
        const uint64_t db = umma_desc(smem_u32(sK + c * chunk_bytes) + k * 32, 0, sbo, p.layout);
        umma_bf16(tmem_s, da, db, idesc, kk > 0 ? 1u : 0u);
      }
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();

  // ---- softmax: thread t owns row t (TMEM lane t) ----
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  float mx = -INFINITY;
  for (int c0 = 0; c0 < kT; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_s + lane_addr + c0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
  }
  const float mxs = mx * p.scale_log2e;
  float sum = 0.f;
  for (int c0 = 0; c0 < kT; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_s + lane_addr + c0, v);
    uint8_t* prow = sP + (c0 >> 6) * (kT * 128) + tid * 128;   // chunk of 64 keys, row = query
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float e = exp2f(fmaf(__uint_as_float(v[j + i]), p.scale_log2e, -mxs));
        const __nv_bfloat16 eb = __float2bfloat16_rn(e);
        f[i] = __bfloat162float(eb);
        sum += f[i];                                            // normalise by what the MMA will see
      }
      const int ch = ((c0 & 63) + j) >> 3;                      // 16 B chunk within the 128 B row
      *reinterpret_cast<bf16x8*>(prow + ((ch ^ (tid & 7)) << 4)) = pack8(f);
    }
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- O = P V : A = P (K-major, SW128, two 64-key chunks), B = V (MN-major as loaded) ----
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_ex(128, static_cast<uint32_t>(p.dk), 0, 1);
    const uint32_t sbo_v = 8u * row_bytes;                      // next 8 keys
    const uint32_t lbo_v = static_cast<uint32_t>(chunk_bytes);  // next 64 head-dim elements (dk = 128 only)
    for (int k = 0; k < kT / 16; ++k) {
      const uint64_t da = umma_desc(smem_u32(sP + (k >> 2) * (kT * 128)) + (k & 3) * 32, 0, 1024, 2);
      const uint64_t db = umma_desc(smem_u32(sV) + k * 16 * row_bytes, lbo_v, sbo_v, p.layout);
      umma_bf16(tmem_o, da, db, idesc, k > 0 ? 1u : 0u);
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 1);
  tc_fence_after();

  // ---- epilogue: out[b*128 + t][h*dk + c] = O[t][c] / sum ----
  const float inv = 1.0f / sum;
  __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * kT + tid) * p.C + h * p.dk;
  for (int c0 = 0; c0 < p.dk; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_o + lane_addr + c0, v);
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (c0 + j < p.dk) {
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[j + i]) * inv;
        *reinterpret_cast<bf16x8*>(orow + c0 + j) = pack8(f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
bool g_attr = false;

}  // namespace

namespace cft {
// Returns CFT_E_UNSUPPORTED when the shape is outside this kernel (caller falls back to the CUDA-core kernel).
int attention_tcgen05(const void* qkv, void* out, int B, int T, int C, int heads, cudaStream_t stream) {
  const int dk = C / heads;
  if (T != kT || C % heads || !(dk == 16 || dk == 32 || dk == 64 || dk == 128) || B > 65535) return CFT_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(qkv) % 16 || (3 * C) % 8) return CFT_E_UNSUPPORTED;
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CFT_E_CUDA;
  }
  AttnParams p;
  p.C = C; p.heads = heads; p.dk = dk;
  p.bw = dk < 64 ? dk : 64;
  p.nchunk = dk / p.bw;
  p.layout = p.bw == 64 ? 2 : (p.bw == 32 ? 4 : 6);
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)(3 * C), (cuuint64_t)B * kT};
  cuuint64_t str[1] = {(cuuint64_t)(3 * C) * 2};
  cuuint32_t box[2] = {(cuuint32_t)p.bw, (cuuint32_t)kT};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle swz = p.bw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                            : (p.bw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(qkv), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(qkv) failed (CUresult %d)", (int)r);
    return CFT_E_CUDA;
  }
  const int smem = 1024 + 3 * kT * dk * 2 + (dk >= 64 ? 0 : 2 * kT * 128) + 64;
  if (!g_attr) {
    int rc = check_cuda(cudaFuncSetAttribute(cft_attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             1024 + 3 * kT * 128 * 2 + 2 * kT * 128 + 64),
                        "cudaFuncSetAttribute(attention_tcgen05)");
    if (rc) return rc;
    g_attr = true;
  }
  dim3 grid(heads, B);
  LaunchScope ls(CFT_K_ATTENTION, stream);
  cft::launch(cft_attention_tcgen05_kernel, dim3(grid), dim3(kThreads), smem, stream, map, p);
  return ls.finish("cft_attention (tcgen05) launch");
}
}  // namespace cft
