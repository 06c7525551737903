// 128-token multi-head self-attention core on the tensor cores (sm_100a), one CTA per (image, head):
//
//   S = Q K^T      tcgen05.mma  128 x 128 x dk   (A = Q, B = K, both K-major smem tiles loaded by TMA
//                                                 straight out of the fused QKV activation [B*128, 3C])
//   P = softmax(S / sqrt(dk))   256 threads: thread t owns query row t % 128 (= TMEM lane) and the key half t / 128
//                               (warps w and w + 4 read the same lane quarter); row max / sum of the two halves are
//                               exchanged through smem; fp32; exp2f with the scale folded in; P written to smem as
//                               the bf16 K-major A operand (each half writes its own 64-key chunk)
//   O = P V        tcgen05.mma  128 x dk x 128   (B = V as loaded: [key][dk] = MN-major operand)
//   out = O / rowsum  -> bf16, heads merged ([B*128, C])
//
// The head dim is cut into chunks of 64 / 32 / 16 elements (widest first: 128 = 64+64, 80 = 64+16, 160 = 64+64+32,
// 40 = 32+16 with the last 8 columns zero-filled by TMA: the tensor maps are 3-D (head dim, q|k|v x head, token), so a box
// that overhangs the head dim reads zeros, not the next head);
// every chunk is its own [128 x width] tile with the swizzle mode that matches its row length (128B / 64B / 32B),
// loaded through the tensor map of that width.  QK^T walks the chunks along K; PV issues one MMA group per chunk
// (N = chunk width) into that chunk's TMEM columns (a uniform 64-wide split is issued as ONE N = dk group).  Replaces models/common.py:497-510 (two batched matmuls + softmax that
// materialise the [B,8,128,128] attention tensor in HBM).
#include "cft_common.cuh"
#include "tcgen05_ptx.cuh"

namespace {
using namespace cft;
using namespace cft::ptx;

constexpr int kT = 128;          // tokens
constexpr int kThreads = 256;

constexpr int kMaxChunks = 4;
struct AttnParams {
  int C, heads, dk;
  int dkp;                    // head dim padded to the chunk widths (multiple of 16): smem tiles / TMEM columns
  int nchunk;                 // head-dim chunks
  int cw[kMaxChunks];         // chunk width in elements (64 / 32 / 16)
  int coff[kMaxChunks];       // first head-dim element of the chunk
  int soff[kMaxChunks];       // byte offset of the chunk's [128 x cw] tile inside a Q / K / V tile
  int layout[kMaxChunks];     // UMMA layout type (2 = SW128, 4 = SW64, 6 = SW32)
  int map[kMaxChunks];        // tensor map index (0: 64-wide, 1: 32-wide, 2: 16-wide boxes)
  int uniform64;              // every chunk is 64 wide -> PV as one N = dk MMA group (LBO = chunk stride)
  int tmem_cols;              // 256 or 512
  float scale_log2e;
  __nv_bfloat16* out;
};
struct __align__(64) AttnMaps {
  CUtensorMap m[3];
};

__global__ void __launch_bounds__(kThreads, 2)
cft_attention_tcgen05_kernel(const __grid_constant__ AttnMaps maps, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int tile_bytes = kT * p.dkp * 2;             // one of Q / K / V (all chunks, head dim zero-padded to dkp)
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
  float* red = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 64);   // [max | sum][key half][row]
  const int row = tid & (kT - 1), half = tid >> 7;

  if (tid == 0) {
    for (int c = 0; c < p.nchunk; ++c) prefetch_tmap(&maps.m[p.map[c]]);
    mbar_init(tma_bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, static_cast<uint32_t>(p.tmem_cols));
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
      const CUtensorMap* m = &maps.m[p.map[c]];     // (head-dim element, part * heads + head, token row)
      tma_load_3d(sQ + p.soff[c], m, tma_bar, p.coff[c], h, r0);
      tma_load_3d(sK + p.soff[c], m, tma_bar, p.coff[c], p.heads + h, r0);
      tma_load_3d(sV + p.soff[c], m, tma_bar, p.coff[c], 2 * p.heads + h, r0);
    }
  }
  mbar_wait(tma_bar, 0);
  tc_fence_after();

  // ---- S = Q K^T : K-major operands; 8-row groups are 8*row_bytes apart (SBO) ----
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_ex(128, 128, 0, 0);
    int kk = 0;
    for (int c = 0; c < p.nchunk; ++c) {
      const uint32_t sbo = 8u * static_cast<uint32_t>(p.cw[c]) * 2u;      // 8-row groups are 8 * row_bytes apart
      for (int k = 0; k < p.cw[c] / 16; ++k, ++kk) {
        const uint64_t da = umma_desc(smem_u32(sQ + p.soff[c]) + k * 32, 0, sbo, p.layout[c]);
        const uint64_t db = umma_desc(smem_u32(sK + p.soff[c]) + k * 32, 0, sbo, p.layout[c]);
        umma_bf16(tmem_s, da, db, idesc, kk > 0 ? 1u : 0u);
      }
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 0);
  tc_fence_after();

  // ---- softmax: thread (row, half) owns 64 of the 128 scores of its row; the scores stay in registers ----
  const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
  uint32_t v0[32], v1[32];
  tmem_ld32(tmem_s + lane_addr + static_cast<uint32_t>(64 * half), v0);
  tmem_ld32(tmem_s + lane_addr + static_cast<uint32_t>(64 * half + 32), v1);
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
  red[half * kT + row] = mx;
  __syncthreads();
  mx = fmaxf(red[row], red[kT + row]);
  const float mxs = mx * p.scale_log2e;
  float sum = 0.f;
  uint8_t* prow = sP + half * (kT * 128) + row * 128;           // chunk of 64 keys, row = query
#pragma unroll
  for (int j = 0; j < 64; j += 8) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t s = j < 32 ? v0[(j + i) & 31] : v1[(j + i) & 31];
      const float e = exp2f(fmaf(__uint_as_float(s), p.scale_log2e, -mxs));
      const __nv_bfloat16 eb = __float2bfloat16_rn(e);
      f[i] = __bfloat162float(eb);
      sum += f[i];                                              // normalise by what the MMA will see
    }
    *reinterpret_cast<bf16x8*>(prow + (((j >> 3) ^ (row & 7)) << 4)) = pack8(f);   // 16 B chunk within the 128 B row
  }
  red[2 * kT + half * kT + row] = sum;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  sum = red[2 * kT + row] + red[3 * kT + row];

  // ---- O = P V : A = P (K-major, SW128, two 64-key chunks), B = V (MN-major as loaded) ----
  if (tid == 0) {
    // one MMA group per head-dim chunk (N = chunk width, its own TMEM columns); a uniform 64-wide split runs as a
    // single N = dk group whose descriptor steps from chunk to chunk through LBO
    const int groups = p.uniform64 ? 1 : p.nchunk;
    for (int c = 0; c < groups; ++c) {
      const uint32_t n = p.uniform64 ? static_cast<uint32_t>(p.dkp) : static_cast<uint32_t>(p.cw[c]);
      const uint32_t row_b = static_cast<uint32_t>(p.cw[c]) * 2u;
      const uint32_t idesc = umma_idesc_ex(128, n, 0, 1);
      const uint32_t sbo_v = 8u * row_b;                        // next 8 keys
      const uint32_t lbo_v = static_cast<uint32_t>(kT) * row_b; // next chunk of head-dim elements (uniform64 only)
      for (int k = 0; k < kT / 16; ++k) {
        const uint64_t da = umma_desc(smem_u32(sP + (k >> 2) * (kT * 128)) + (k & 3) * 32, 0, 1024, 2);
        const uint64_t db = umma_desc(smem_u32(sV + p.soff[c]) + k * 16 * row_b, lbo_v, sbo_v, p.layout[c]);
        umma_bf16(tmem_o + static_cast<uint32_t>(p.coff[c]), da, db, idesc, k > 0 ? 1u : 0u);
      }
    }
    umma_commit(mma_bar);
  }
  mbar_wait(mma_bar, 1);
  tc_fence_after();

  // ---- epilogue: out[b*128 + t][h*dk + c] = O[t][c] / sum; the two halves take alternate 32-column groups ----
  const float inv = 1.0f / sum;
  __nv_bfloat16* orow = p.out + (static_cast<size_t>(b) * kT + row) * p.C + h * p.dk;
  for (int c0 = 32 * half; c0 < p.dkp; c0 += 64) {
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
  if (warp == 0) tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
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

int launch_attention(const AttnMaps& maps, const AttnParams& p, int B, cudaStream_t stream) {
  const int heads = p.heads;
  const int dk = p.dkp;
  const int smem = 1024 + 3 * kT * dk * 2 + (dk >= 64 ? 0 : 2 * kT * 128) + 64 + 4 * kT * 4;
  if (!g_attr) {
    int rc = check_cuda(cudaFuncSetAttribute(cft_attention_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             1024 + 3 * kT * 256 * 2 + 64 + 4 * kT * 4),
                        "cudaFuncSetAttribute(attention_tcgen05)");
    if (rc) return rc;
    g_attr = true;
  }
  dim3 grid(heads, B);
  LaunchScope ls(CFT_K_ATTENTION, stream);
  cft::launch(cft_attention_tcgen05_kernel, dim3(grid), dim3(kThreads), smem, stream, maps, p);
  return ls.finish("cft_attention (tcgen05) launch");
}
}  // namespace

namespace cft {
// Returns CFT_E_UNSUPPORTED when the shape is outside this kernel (caller falls back to the CUDA-core kernel).
int attention_tcgen05(const void* qkv, void* out, int B, int T, int C, int heads, cudaStream_t stream) {
  const int dk = C / heads;
  if (T != kT || C % heads || dk % 8 || dk < 16 || dk > 256 || B > 65535) return CFT_E_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(qkv) % 16 || (3 * C) % 8) return CFT_E_UNSUPPORTED;
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CFT_E_CUDA;
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.C = C; p.heads = heads; p.dk = dk;
  {
    int off = 0, soff = 0, n = 0;
    bool used[3] = {false, false, false};
    while (off < dk) {
      const int rem = dk - off;
      const int w = rem >= 64 ? 64 : (rem >= 32 ? 32 : 16);     // the last chunk may overhang (rem = 8): zero-filled
      if (n == kMaxChunks) return CFT_E_UNSUPPORTED;
      p.cw[n] = w;
      p.coff[n] = off;
      p.soff[n] = soff;
      p.layout[n] = w == 64 ? 2 : (w == 32 ? 4 : 6);
      p.map[n] = w == 64 ? 0 : (w == 32 ? 1 : 2);
      used[p.map[n]] = true;
      off += w;
      soff += kT * w * 2;
      ++n;
    }
    p.nchunk = n;
    p.dkp = off;
    p.uniform64 = (dk % 64 == 0) ? 1 : 0;
    p.tmem_cols = (128 + p.dkp <= 256) ? 256 : 512;
    p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    AttnMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int i = 0; i < 3; ++i) {
      if (!used[i]) continue;
      const int w = i == 0 ? 64 : (i == 1 ? 32 : 16);
      // qkv [B*128, 3C] viewed as (head-dim element, q|k|v x head, token): a box overhanging dk is zero-filled
      cuuint64_t dims[3] = {(cuuint64_t)dk, (cuuint64_t)(3 * heads), (cuuint64_t)B * kT};
      cuuint64_t str[2] = {(cuuint64_t)dk * 2, (cuuint64_t)(3 * C) * 2};
      cuuint32_t box[3] = {(cuuint32_t)w, 1, (cuuint32_t)kT};
      cuuint32_t estr[3] = {1, 1, 1};
      const CUtensorMapSwizzle swz = w == 64 ? CU_TENSOR_MAP_SWIZZLE_128B
                                             : (w == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
      CUresult r = enc(&maps.m[i], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), dims, str, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(qkv, box %d) failed (CUresult %d)", w, (int)r);
        return CFT_E_CUDA;
      }
    }
    return launch_attention(maps, p, B, stream);
  }
}

}  // namespace cft
