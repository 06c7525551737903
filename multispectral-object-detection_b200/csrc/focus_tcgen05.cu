// Fused Focus layer (reference models/common.py:168-180, fed by the loader's uint8 wire format,
// utils/datasets.py:1272-1281 + train.py:715 / test.py:107-108): space-to-depth + concat + 3x3 conv + BN + SiLU
// in ONE kernel, straight from the uint8 image.
//
// Space-to-depth followed by a 3x3 / pad 1 conv is a 6x6 / stride 2 / pad 2 conv on the 3-channel image:
//   out(oy, ox, o) = sum_{c, r, q} img[c, 2 oy + r - 2, 2 ox + q - 2] * Wf[o, (c, r, q)],  r, q in 0..5,
//   Wf[o, (c, r = 2 ky + gy, q = 2 kx + gx)] = W[o, (gy + 2 gx) * 3 + c, ky, kx]       (packed by the host side).
// Per 16 x 8 output tile:
//   warp 0      TMA: the (64 x 20 x 3)-byte image patch (out-of-image bytes arrive as zeros = the conv padding)
//   warps 4-11  build the im2col A tile [128 pixels x K = 144] in smem, already in the 128B-swizzled K-major UMMA
//               layout.  K = 18 chunks (c, r) of 8: six horizontally adjacent image bytes + 2 zeros.  The bytes are
//               converted to fp16 EXACTLY (PRMT to 0x64xx = 1024 + x, minus 1024), so the only rounding of this layer
//               is that of the fp16 weights; 1/255 is applied to the fp32 accumulator.
//   warp 1      9 tcgen05.mma (kind::f16, fp16 x fp16 -> fp32 in TMEM) per tile, weights resident in smem
//   warps 12-19 epilogue: tcgen05.ld -> acc / 255 + bias -> SiLU -> bf16 -> swizzled staging -> TMA store (NHWC slice)
// Replaces cft_focus_gather + the 16-channel cft_conv2d whose 32-byte operand rows ran at 14 % of the tensor peak.
#include <cuda_fp16.h>

#include "cft_common.cuh"
#include "tcgen05_ptx.cuh"

namespace cft {
namespace {

using namespace cft::ptx;

constexpr int kFThreads = 640;
constexpr int kFEpiWarps = 8;             // warps 12..19: two per TMEM lane quarter, alternating 32-column chunks
constexpr int kFBuilderWarps = 8;         // warps 4..11
constexpr int kFTileW = 16, kFTileH = 8;  // output pixels per tile (128 = one UMMA M)
constexpr int kFPatchW = 64, kFPatchH = 20;                 // image bytes [2 x0 - 16, 2 x0 + 48) x [2 y0 - 2, 2 y0 + 18): the
                                                            // 36 needed columns start at byte 14 (TMA boxes start on 16 B)
constexpr int kFPatchX0 = 14;
constexpr int kFPatchBytes = kFPatchW * kFPatchH * 3;       // 3840
constexpr int kFPatchSlot = 3840;                           // 128 B aligned
constexpr int kFMaxPatches = 6;                             // TMA latency / 6 in flight stays below the per-tile time
constexpr int kFMaxAStages = 3;
constexpr int kFAtomBytes = 128 * 128;                      // 128 rows x 64 fp16
constexpr int kFATileBytes = 3 * kFAtomBytes;               // K = 144 -> atoms 0, 1 full, atom 2: one 16-element step
constexpr int kFMaxCout = 128;
constexpr int kFStageC = 8 * 1024;                          // one 32-channel bf16 chunk of a tile (128 rows x 64 B)
constexpr int kFAccStages = 4;                              // 4 x 128 TMEM columns

struct FocusMaps {
  CUtensorMap img;   // uint8 (W, H, 3, B), box (64, 20, 3, 1), no swizzle
  CUtensorMap w;     // fp16 [Cout][192], box (64, Cout), SWIZZLE_128B
  CUtensorMap c;     // bf16 out (Cout, Wo, Ho, B), box (32, 16, 2, 1), SWIZZLE_64B
};

struct FocusParams {
  int B, Ho, Wo, Cout;
  int tiles_x, tiles_y, num_tiles;
  int act;              // CFT_ACT_NONE / CFT_ACT_SILU
  int chunks;           // ceil(Cout / 32)
  int patches;          // image-patch ring depth (<= kFMaxPatches)
  int a_stages;         // A-tile ring depth (3, or 2 when the weights / staging of a wide layer need the room)
  const float* bias;
};

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

__global__ void __launch_bounds__(kFThreads, 1)
cft_focus_tcgen05_kernel(const __grid_constant__ FocusMaps maps, const __grid_constant__ FocusParams p) {
  extern __shared__ uint8_t fsmem_raw[];
  uint8_t* smem = fsmem_raw + ((1024u - (smem_u32(fsmem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  uint8_t* smem_a = smem;                                              // [a_stages][3 atoms][128 rows][128 B]
  uint8_t* smem_w = smem_a + p.a_stages * kFATileBytes;                // [3 atoms][Cout rows][128 B]
  uint8_t* smem_c = smem_w + 3 * p.Cout * 128;                         // [2][chunks][8 KiB]
  uint8_t* smem_p = smem_c + 2 * p.chunks * kFStageC;                  // [kFMaxPatches][3072]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + p.patches * kFPatchSlot);
  uint64_t* pfull = bars;                    // [kFMaxPatches]  TMA -> builders
  uint64_t* pempty = bars + kFMaxPatches;       // [kFMaxPatches]  builders -> TMA
  uint64_t* afull = bars + 2 * kFMaxPatches;    // [kFMaxAStages]  builders -> MMA
  uint64_t* aempty = afull + kFMaxAStages;      // [kFMaxAStages]  MMA -> builders
  uint64_t* tfull = aempty + kFMaxAStages;      // [kFAccStages] MMA -> epilogue
  uint64_t* tempty = tfull + kFAccStages;    // [kFAccStages] epilogue -> MMA
  uint64_t* wbar = tempty + kFAccStages;     // weights landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [128], pre-scaled for the activation

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.img);
    prefetch_tmap(&maps.w);
    prefetch_tmap(&maps.c);
    for (int i = 0; i < kFMaxPatches; ++i) {
      mbar_init(&pfull[i], 1);
      mbar_init(&pempty[i], kFBuilderWarps);
    }
    for (int i = 0; i < kFMaxAStages; ++i) {
      mbar_init(&afull[i], kFBuilderWarps);
      mbar_init(&aempty[i], 1);
    }
    for (int i = 0; i < kFAccStages; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kFEpiWarps);
    }
    mbar_init(wbar, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < kFMaxCout; i += kFThreads) {
    const float bv = (p.bias != nullptr && i < p.Cout) ? __ldg(p.bias + i) : 0.f;
    bias_s[i] = p.act == CFT_ACT_SILU ? 0.5f * bv : bv;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  const int tiles_per_img = p.tiles_x * p.tiles_y;
  auto decode = [&](int tile, int& b, int& y0, int& x0) {
    b = tile / tiles_per_img;
    const int r = tile - b * tiles_per_img;
    const int ty = r / p.tiles_x;
    y0 = ty * kFTileH;
    x0 = (r - ty * p.tiles_x) * kFTileW;
  };

  if (warp == 0) {
    // ===================== TMA: weights once, then one image patch per tile =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(wbar, 3u * static_cast<uint32_t>(p.Cout) * 128u);
      for (int a = 0; a < 3; ++a) tma_load_2d(smem_w + a * p.Cout * 128, &maps.w, wbar, a * 64, 0);
    }
    int slot = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int b, y0, x0;
      decode(tile, b, y0, x0);
      mbar_wait(&pempty[slot], phase ^ 1u);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&pfull[slot], kFPatchBytes);
        tma_load_4d(smem_p + slot * kFPatchSlot, &maps.img, &pfull[slot], 2 * x0 - 2 - kFPatchX0, 2 * y0 - 2, 0, b);
      }
      __syncwarp();
      if (++slot == p.patches) {
        slot = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(p.Cout) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);      // SBO = 1024 B, version 1, SWIZZLE_128B
    const uint32_t a_lo0 = smem_u32(smem_a) >> 4, w_lo0 = smem_u32(smem_w) >> 4;
    const uint32_t w_atom16 = (static_cast<uint32_t>(p.Cout) * 128u) >> 4;
    mbar_wait(wbar, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty[acc], acc_phase ^ 1u);
      mbar_wait(&afull[stage], phase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 128);
        const uint32_t a_lo = a_lo0 + stage * (kFATileBytes >> 4);
#pragma unroll
        for (int s = 0; s < 9; ++s) {            // K steps of 16: atoms 0, 1 hold four each, atom 2 the last one
          const uint32_t atom = s >> 2, k = s & 3;
          const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + atom * (kFAtomBytes >> 4) + 2 * k);
          const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (w_lo0 + atom * w_atom16 + 2 * k);
          umma_f16(d_tmem, da, db, idesc, s != 0 ? 1u : 0u);
        }
        umma_commit(&aempty[stage]);
        umma_commit(&tfull[acc]);
      }
      __syncwarp();
      if (++stage == p.a_stages) {
        stage = 0;
        phase ^= 1u;
      }
      if (++acc == kFAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4 && warp < 4 + kFBuilderWarps) {
    // ===================== A-tile builders: 2 threads per output pixel =====================
    const int bt = threadIdx.x - 128;          // 0..255
    const int m = bt & 127;                    // pixel (row of the A tile)
    const int half = bt >> 7;                  // image rows r = 3 * half + (0..2) of every channel
    const int oxl = m & 15, oyl = m >> 4;
    const uint32_t b0 = static_cast<uint32_t>(kFPatchX0 + 2 * oxl);   // first of the 6 bytes within the 64-byte patch row
    const uint32_t odd16 = ((b0 >> 1) & 1u) * 16;                       // they start at byte 0 or 2 of an aligned word pair
    const uint32_t col_byte = b0 & ~3u;
    const uint32_t row0 = static_cast<uint32_t>(2 * oyl + 3 * half);
    const uint32_t a_row = static_cast<uint32_t>(m) * 128u;
    const uint32_t sw = static_cast<uint32_t>(m & 7);
    int slot = 0, stage = 0;
    uint32_t pphase = 0, aphase = 0;
    const uint32_t k1024 = 0x64006400u;
    const __half2 h1024 = *reinterpret_cast<const __half2*>(&k1024);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&aempty[stage], aphase ^ 1u);
      mbar_wait(&pfull[slot], pphase);
      const uint8_t* patch = smem_p + slot * kFPatchSlot;
      uint8_t* a_tile = smem_a + stage * kFATileBytes;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const uint32_t prow = static_cast<uint32_t>(c * kFPatchH) + row0 + i;
          const uint32_t* src = reinterpret_cast<const uint32_t*>(patch + prow * kFPatchW + col_byte);
          const uint32_t w0 = src[0], w1 = src[1];
          const uint32_t t0 = __funnelshift_r(w0, w1, odd16);      // bytes d0..d3 of the six
          const uint32_t t1 = w1 >> odd16;                         // bytes d4, d5
          uint32_t q0 = __byte_perm(t0, 0x64646464u, 0x4140);      // {d0, 0x64, d1, 0x64} = fp16 (1024 + d0, 1024 + d1)
          uint32_t q1 = __byte_perm(t0, 0x64646464u, 0x4342);
          uint32_t q2 = __byte_perm(t1, 0x64646464u, 0x4140);
          __half2 h0 = __hsub2(*reinterpret_cast<__half2*>(&q0), h1024);
          __half2 h1 = __hsub2(*reinterpret_cast<__half2*>(&q1), h1024);
          __half2 h2 = __hsub2(*reinterpret_cast<__half2*>(&q2), h1024);
          const uint32_t j = static_cast<uint32_t>(c * 6 + 3 * half + i);     // K chunk (c, r): elements 8 j .. 8 j + 7
          const uint32_t atom = j >> 3, cj = j & 7u;
          uint4 o;
          o.x = *reinterpret_cast<uint32_t*>(&h0);
          o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2);
          o.w = 0u;
          *reinterpret_cast<uint4*>(a_tile + atom * kFAtomBytes + a_row + ((cj ^ sw) << 4)) = o;
        }
      }
      fence_proxy_async();              // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&afull[stage]);
        mbar_arrive(&pempty[slot]);
      }
      if (++slot == p.patches) {
        slot = 0;
        pphase ^= 1u;
      }
      if (++stage == p.a_stages) {
        stage = 0;
        aphase ^= 1u;
      }
    }
  } else if (warp >= 12) {
    // ===================== epilogue (8 warps: lane quarter = warp % 4, chunk parity = (warp - 12) / 4) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int cg = (warp - 12) >> 2;
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    const float scale = p.act == CFT_ACT_SILU ? 0.5f / 255.f : 1.f / 255.f;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int b, y0, x0;
      decode(tile, b, y0, x0);
      uint8_t* stage_c = smem_c + buf * p.chunks * kFStageC;
      if (lane == 0) bulk_wait_read<1>();        // this warp's store that last used the buffer has drained it
      __syncwarp();
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(acc * 128);
      for (int ch = cg; ch < p.chunks; ch += 2) {
        uint32_t v[32];
        tmem_ld32(t_row + static_cast<uint32_t>(ch * 32), v);
        float f[32];
        const float4* bs = reinterpret_cast<const float4*>(bias_s + ch * 32);
        if (p.act == CFT_ACT_SILU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = bs[i];
            f[4 * i + 0] = silu_tanh_h(fmaf(__uint_as_float(v[4 * i + 0]), scale, b4.x));
            f[4 * i + 1] = silu_tanh_h(fmaf(__uint_as_float(v[4 * i + 1]), scale, b4.y));
            f[4 * i + 2] = silu_tanh_h(fmaf(__uint_as_float(v[4 * i + 2]), scale, b4.z));
            f[4 * i + 3] = silu_tanh_h(fmaf(__uint_as_float(v[4 * i + 3]), scale, b4.w));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b4 = bs[i];
            f[4 * i + 0] = fmaf(__uint_as_float(v[4 * i + 0]), scale, b4.x);
            f[4 * i + 1] = fmaf(__uint_as_float(v[4 * i + 1]), scale, b4.y);
            f[4 * i + 2] = fmaf(__uint_as_float(v[4 * i + 2]), scale, b4.z);
            f[4 * i + 3] = fmaf(__uint_as_float(v[4 * i + 3]), scale, b4.w);
          }
        }
        uint8_t* st = stage_c + ch * kFStageC;     // TMA box (32 ch, 16, 8): 64 B rows, SWIZZLE_64B
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4)
          *reinterpret_cast<bf16x8*>(st + row * 64 + ((c4 ^ ((row >> 1) & 3)) << 4)) = pack8(f + 8 * c4);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      fence_proxy_async();
      __syncwarp();
      // every warp stores its own 32 accumulator rows = two 16-pixel rows of the tile: no cross-warp barrier in the loop
      if (lane == 0) {
        for (int ch = cg; ch < p.chunks; ch += 2)
          tma_store_4d(&maps.c, stage_c + ch * kFStageC + q * 2048, ch * 32, x0, y0 + 2 * q, b);
        bulk_commit();
      }
      buf ^= 1;
      if (++acc == kFAccStages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode(CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
           const cuuint64_t* strides_b, const cuuint32_t* box, CUtensorMapSwizzle swz) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CFT_E_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(m, dt, rank, const_cast<void*>(base), dims, strides_b, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cft_focus_conv: cuTensorMapEncodeTiled failed (CUresult %d, rank %d)", (int)r, rank);
    return CFT_E_CUDA;
  }
  return CFT_OK;
}

bool g_focus_attr = false;

}  // namespace
}  // namespace cft

using namespace cft;

extern "C" int cft_focus_conv(const void* img, int B, int H, int W, long long batch_stride, const void* w, const float* bias,
                              int Cout, int act, void* y, int ldy, int y_coff, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CFT_REQUIRE(img && w && y, "cft_focus_conv: null pointer");
  CFT_REQUIRE(B > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "cft_focus_conv: H, W must be even (got %d x %d)", H, W);
  CFT_REQUIRE(W % 16 == 0 && batch_stride % 16 == 0 && reinterpret_cast<uintptr_t>(img) % 16 == 0,
              "cft_focus_conv: the uint8 image needs 16-byte aligned rows (W %% 16 == 0) and base");
  CFT_REQUIRE(Cout % 16 == 0 && Cout >= 16 && Cout <= kFMaxCout, "cft_focus_conv: Cout must be a multiple of 16 in [16, 128]");
  CFT_REQUIRE(act == CFT_ACT_NONE || act == CFT_ACT_SILU, "cft_focus_conv: act must be none or SiLU");
  CFT_REQUIRE(ldy % 8 == 0 && y_coff % 8 == 0 && y_coff + Cout <= ldy && reinterpret_cast<uintptr_t>(y) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(w) % 16 == 0,
              "cft_focus_conv: output slice / weight alignment");

  FocusParams p;
  p.B = B;
  p.Ho = H / 2;
  p.Wo = W / 2;
  p.Cout = Cout;
  p.tiles_x = (p.Wo + kFTileW - 1) / kFTileW;
  p.tiles_y = (p.Ho + kFTileH - 1) / kFTileH;
  const long long nt = static_cast<long long>(B) * p.tiles_x * p.tiles_y;
  CFT_REQUIRE(nt < (1LL << 31), "cft_focus_conv: too many tiles");
  p.num_tiles = static_cast<int>(nt);
  p.act = act;
  p.chunks = (Cout + 31) / 32;
  p.bias = bias;

  FocusMaps maps;
  int rc;
  {
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W, (cuuint64_t)H * W, (cuuint64_t)batch_stride};
    cuuint32_t box[4] = {kFPatchW, kFPatchH, 3, 1};
    rc = encode(&maps.img, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, img, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[2] = {192, (cuuint64_t)Cout};
    cuuint64_t str[1] = {192 * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)Cout};
    rc = encode(&maps.w, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, w, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    const uint8_t* yb = reinterpret_cast<const uint8_t*>(y) + static_cast<size_t>(y_coff) * 2;
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)p.Wo, (cuuint64_t)p.Ho, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)ldy * 2, (cuuint64_t)p.Wo * ldy * 2, (cuuint64_t)p.Ho * p.Wo * ldy * 2};
    cuuint32_t box[4] = {32, kFTileW, 2, 1};        // one epilogue warp's share of a tile: 32 pixels = 2 rows
    rc = encode(&maps.c, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, yb, dims, str, box, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
  }

  p.patches = kFMaxPatches;
  p.a_stages = kFMaxAStages;
  auto smem_need = [&]() {
    return 1024 + p.a_stages * kFATileBytes + 3 * Cout * 128 + 2 * p.chunks * kFStageC + p.patches * kFPatchSlot + 256 + kFMaxCout * 4;
  };
  while (smem_need() > 227 * 1024) {        // wide layers: give up ring depth for their weights and staging
    if (p.a_stages > 2) --p.a_stages;
    else if (p.patches > 3) --p.patches;
    else break;
  }
  const int smem_bytes = smem_need();
  CFT_REQUIRE(smem_bytes <= 227 * 1024, "cft_focus_conv: shared memory budget exceeded (Cout %d)", Cout);
  if (!g_focus_attr) {
    rc = check_cuda(cudaFuncSetAttribute(cft_focus_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024),
                    "cudaFuncSetAttribute(focus_tcgen05)");
    if (rc) return rc;
    g_focus_attr = true;
  }
  LaunchScope ls(CFT_K_FOCUS, stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p.num_tiles < sm_count() ? p.num_tiles : sm_count());
  cfg.blockDim = dim3(kFThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, cft_focus_tcgen05_kernel, maps, p);
  if (e != cudaSuccess) {
    ls.finish("cft_focus_conv launch");
    return check_cuda(e, "cudaLaunchKernelEx(focus_tcgen05)");
  }
  return ls.finish("cft_focus_conv launch");
}
