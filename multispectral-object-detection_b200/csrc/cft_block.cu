// The whole transformer stack of one CFT / GPT block (models/common.py:593-639, the n_layer x myTransformerBlock loop
// at :622 plus ln_f at :625; block = :516-546, attention = :475-513) as ONE kernel launch (sm_100a):
//
//   for every layer:  x += out_proj(MHSA(LN1(x)));   x += W2 GELU(W1 LN2(x) + b1) + b2;     finally  y = ln_f(x)
//
// One thread-block CLUSTER of C CTAs owns one image (its 128 x d token tile) for all layers; images are independent,
// so there is no grid-wide synchronisation, only cluster barriers.  CTA `c` of a cluster owns the DC = d / C columns
// [c*DC, (c+1)*DC) of the residual stream (fp32, in REGISTERS for the whole kernel), the heads that live in those
// columns, and the matching N-slices of all four Linear layers:
//
//   QKV   A = LN1(x) [128, d] resident in smem; B = this CTA's q|k|v weight rows streamed by TMA; the accumulator is
//         drained (+bias, bf16) straight into UMMA-layout Q / K / V smem tiles -> attention never leaves the SM
//   attn  S = Q K^T (tcgen05), softmax by 256 threads out of TMEM, P -> smem, O = P V (tcgen05), O / rowsum -> L2
//   out   A = O (all heads, all-gathered through L2) resident; N = DC; epilogue: x += acc + bias, LN2 statistics
//         exchanged between the CTAs through distributed shared memory, LN2(x) slice -> L2
//   up    A = LN2(x) resident; N = 4 DC in passes of 256; epilogue: +bias, erf-GELU, bf16 hidden slice -> L2
//   down  A = hidden [128, 4d] streamed through the operand slots; N = DC; epilogue: x += acc + bias, next LN1 / ln_f
//
// Warp roles (384 threads): warp 0 weight (B) producer | warp 1 MMA issuer | warp 2 activation (A) producer + TMEM
// allocator | warp 3 idle (keeps the cluster-barrier count) | warps 4-11 compute: thread (t, hh) owns token row
// t = TMEM lane t and half hh of every column range.
// Six cluster barriers per layer publish the all-gathered operands (O, LN2(x), hidden, LN1(x)) and the LayerNorm
// partial sums; the weight producer and the MMA issuer use the split arrive / wait form so that weights of the next
// GEMM are prefetched across a barrier.
//
// Replaces 56 dependent launches per block (7 per layer: LN, QKV GEMM, attention, out-proj, LN, MLP up, MLP down).
#include <stdlib.h>

#include "cft_common.cuh"
#include "tcgen05_ptx.cuh"

namespace {
using namespace cft;
using namespace cft::ptx;

constexpr int kT = 128;                    // tokens per image (2 * 8 * 8)
constexpr int kComputeWarps = 8;
constexpr int kThreads = 128 + 32 * kComputeWarps;   // 384
constexpr int kStageBytes = 16384;         // weight ring stage: <= 256 rows x 32 k (64 B rows, SWIZZLE_64B)
constexpr int kAChunk = 16384;             // activation chunk: 128 rows x 64 k (128 B rows, SWIZZLE_128B)
constexpr int kMaxStages = 8;
constexpr int kMaxASlots = 8;
constexpr int kTmemCols = 512;
constexpr uint32_t kTmemS = 384;           // S = Q K^T accumulator columns [384, 512)
constexpr int kSmemMax = 227 * 1024;

struct __align__(64) BlockMaps {
  CUtensorMap wqkv, wo, w1, w2;   // weights, box {32 k, rows}
  CUtensorMap abuf, hbuf;         // activations, box {64 k, 128 rows}
};

struct BlockParams {
  int B, d, heads, dk, layers, C, hpc;
  int stages, a_slots, ra_bytes;
  int cw, nch, layout, rowB;      // attention tiles: chunk width (elements), chunks per head, UMMA layout code, row bytes
  float scale_log2e, eps1, eps2, epsf;
  const float *bqkv, *bo, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b, *lnfg, *lnfb;
  const float* x_in;
  float* x_out;
  __nv_bfloat16* abuf;            // [3][B*128][d]: LN1(x) | O | LN2(x)
  __nv_bfloat16* hbuf;            // [B*128][4d]
  float* dbg;                     // optional [layers][B][128][d] dump of x after every layer
};

struct Ring {
  int stage;
  uint32_t phase;
};

__device__ __forceinline__ void ring_advance(Ring& r, int stages) {
  if (++r.stage == stages) {
    r.stage = 0;
    r.phase ^= 1u;
  }
}

// ------------------------------------------------------------------ weight producer: one GEMM pass
__device__ __forceinline__ void produce_pass(const CUtensorMap* map, int nseg, int row0, int row1, int row2, int seg_rows,
                                             int k32, uint8_t* ring, uint64_t* bfull, uint64_t* bempty, int stages,
                                             Ring& r) {
  const uint32_t tx = static_cast<uint32_t>(nseg * seg_rows) * 64u;
  for (int i = 0; i < k32; ++i) {
    mbar_wait(&bempty[r.stage], r.phase ^ 1u);
    if (elect_one_sync()) {
      uint8_t* dst = ring + r.stage * kStageBytes;
      mbar_arrive_expect_tx(&bfull[r.stage], tx);
      tma_load_2d(dst, map, &bfull[r.stage], i * 32, row0);
      if (nseg > 1) tma_load_2d(dst + seg_rows * 64, map, &bfull[r.stage], i * 32, row1);
      if (nseg > 2) tma_load_2d(dst + 2 * seg_rows * 64, map, &bfull[r.stage], i * 32, row2);
    }
    __syncwarp();
    ring_advance(r, stages);
  }
}

// ------------------------------------------------------------------ MMA issuer: one GEMM pass
// a_mode 0: A resident, chunk barriers not yet consumed | 1: A resident, already waited for | 2: A streamed (slot ring)
__device__ __forceinline__ void mma_pass(uint32_t ra_addr, uint32_t ring_addr, int a_mode, int n, int k32,
                                         uint32_t d_tmem, uint64_t* afull, uint64_t* aempty, int a_slots,
                                         uint32_t& a_par, uint64_t* bfull, uint64_t* bempty, int stages, Ring& r,
                                         uint64_t* acc_bar) {
  const uint32_t idesc = umma_idesc_ex(128u, static_cast<uint32_t>(n), 0, 0);
  constexpr uint32_t a_hi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO 1024 B, SWIZZLE_128B
  constexpr uint32_t b_hi = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO 512 B, SWIZZLE_64B
  for (int i = 0; i < k32; ++i) {
    const int chunk = i >> 1;
    const int slot = a_mode == 2 ? chunk % a_slots : chunk;
    if ((i & 1) == 0 && a_mode != 1) {
      mbar_wait(&afull[slot], (a_par >> slot) & 1u);
      a_par ^= 1u << slot;
    }
    mbar_wait(&bfull[r.stage], r.phase);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint32_t a_lo = (ra_addr + static_cast<uint32_t>(slot) * kAChunk + static_cast<uint32_t>(i & 1) * 64u) >> 4;
      const uint32_t b_lo = (ring_addr + static_cast<uint32_t>(r.stage) * kStageBytes) >> 4;
#pragma unroll
      for (int s = 0; s < 2; ++s) {            // two K = 16 steps per 32-element stage (+32 B inside the swizzle atom)
        const uint64_t da = (static_cast<uint64_t>(a_hi) << 32) | ((a_lo + 2u * s) & 0x3FFFu);
        const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | ((b_lo + 2u * s) & 0x3FFFu);
        umma_bf16(d_tmem, da, db, idesc, (i | s) != 0 ? 1u : 0u);
      }
      umma_commit(&bempty[r.stage]);
      if (a_mode == 2 && (i & 1)) umma_commit(&aempty[slot]);
      if (i == k32 - 1) umma_commit(acc_bar);
    }
    __syncwarp();
    ring_advance(r, stages);
  }
}

// ------------------------------------------------------------------ the kernel
template <int DC>
__global__ void __launch_bounds__(kThreads, 1)
cft_gpt_block_kernel(const __grid_constant__ BlockMaps maps, const __grid_constant__ BlockParams p) {
  constexpr int NC = DC / 2;                       // residual-stream columns per compute thread
  constexpr int NQP = (3 * DC <= 256) ? 1 : 2;     // QKV passes: q|k|v (N = 192) or q|k (256) + v (128)
  constexpr int NUP = (4 * DC) / 256;              // MLP-up passes of N = 256
  constexpr int TB = kT * DC * 2;                  // bytes of this CTA's Q (or K, or V) tiles, all local heads

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* RA = smem;                                          // resident A operand | Q,K,V,P tiles | streamed A slots
  uint8_t* ring = RA + p.ra_bytes;                             // weight stages
  uint8_t* misc = ring + p.stages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint64_t* bfull = bars;                                      // [8]
  uint64_t* bempty = bars + kMaxStages;                        // [8]
  uint64_t* afull = bars + 2 * kMaxStages;                     // [8]
  uint64_t* aempty = bars + 2 * kMaxStages + kMaxASlots;       // [8]
  uint64_t* acc_bar = bars + 2 * kMaxStages + 2 * kMaxASlots;  // [2]
  uint64_t* op_bar = acc_bar + 2;                              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(op_bar + 1);
  float* smax = reinterpret_cast<float*>(misc + 512);          // [2][128]
  float* ssum = smax + 2 * kT;                                 // [2][128]
  float2* stats = reinterpret_cast<float2*>(misc + 512 + 2048);   // [2C][128] LayerNorm partial (sum, sum of squares)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = p.C, d = p.d;
  const int rank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x / C, n_clusters = gridDim.x / C;

  if (threadIdx.x == 0) {
    prefetch_tmap(&maps.wqkv);
    prefetch_tmap(&maps.wo);
    prefetch_tmap(&maps.w1);
    prefetch_tmap(&maps.w2);
    prefetch_tmap(&maps.abuf);
    prefetch_tmap(&maps.hbuf);
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < kMaxASlots; ++i) {
      mbar_init(&afull[i], 1);
      mbar_init(&aempty[i], 1);
    }
    mbar_init(&acc_bar[0], 1);
    mbar_init(&acc_bar[1], 1);
    mbar_init(op_bar, kComputeWarps);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  // peers write LayerNorm partials into this CTA's shared memory: nobody may run ahead of a CTA that has not started
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // (no early launch_dependents: a dependent kernel must not take SMs from clusters still to be scheduled)

  const int k32_d = d / 32, kch_d = d / 64;
  const int cpi = 2 + 6 * p.layers;               // cluster barriers per image

  if (warp == 0) {
    // ===================================================== weight (B) producer
    Ring r{0, 0u};
    bool pend = false;
    auto sp = [&]() {
      if (pend) cluster_wait_acquire();
      cluster_arrive_release();
      pend = true;
    };
    for (int b = cluster_id; b < p.B; b += n_clusters) {
      sp();
      sp();
      for (int l = 0; l < p.layers; ++l) {
        const int rq = l * 3 * d + rank * DC;
        if (NQP == 1) {
          produce_pass(&maps.wqkv, 3, rq, rq + d, rq + 2 * d, DC, k32_d, ring, bfull, bempty, p.stages, r);
        } else {
          produce_pass(&maps.wqkv, 2, rq, rq + d, 0, DC, k32_d, ring, bfull, bempty, p.stages, r);
          produce_pass(&maps.wqkv, 1, rq + 2 * d, 0, 0, DC, k32_d, ring, bfull, bempty, p.stages, r);
        }
        sp();                                                                              // #A
        produce_pass(&maps.wo, 1, l * d + rank * DC, 0, 0, DC, k32_d, ring, bfull, bempty, p.stages, r);
        sp();                                                                              // #B
        sp();                                                                              // #C
        for (int u = 0; u < NUP; ++u)
          produce_pass(&maps.w1, 1, l * 4 * d + rank * 4 * DC + u * 256, 0, 0, 256, k32_d, ring, bfull, bempty,
                       p.stages, r);
        sp();                                                                              // #D
        produce_pass(&maps.w2, 1, l * d + rank * DC, 0, 0, DC, 4 * k32_d, ring, bfull, bempty, p.stages, r);
        sp();                                                                              // #E
        sp();                                                                              // #F
      }
    }
    if (pend) cluster_wait_acquire();
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    Ring r{0, 0u};
    uint32_t a_par = 0u;
    uint32_t ev = 0u, opn = 0u;
    bool pend = false;
    auto sp = [&]() {
      if (pend) cluster_wait_acquire();
      cluster_arrive_release();
      pend = true;
    };
    auto acc = [&]() -> uint64_t* { return &acc_bar[(ev++) & 1u]; };
    const uint32_t ra_addr = smem_u32(RA), ring_addr = smem_u32(ring);
    const uint32_t tile_b = static_cast<uint32_t>(kT * p.dk * 2);           // one head's Q (or K, V) tile
    const uint32_t rowB = static_cast<uint32_t>(p.rowB);
    const uint32_t chunk_b = static_cast<uint32_t>(kT) * rowB;
    const uint32_t idesc_s = umma_idesc_ex(128, 128, 0, 0);
    const uint32_t idesc_o = umma_idesc_ex(128, static_cast<uint32_t>(p.dk), 0, 1);
    auto issue_s = [&](int h) {
      if (elect_one_sync()) {
        const uint32_t q0 = ra_addr + static_cast<uint32_t>(h) * tile_b, k0 = q0 + TB;
        int kk = 0;
        for (int ci = 0; ci < p.nch; ++ci)
          for (int k = 0; k < p.cw / 16; ++k, ++kk) {
            const uint64_t da = umma_desc(q0 + ci * chunk_b + k * 32, 0, 8u * rowB, p.layout);
            const uint64_t db = umma_desc(k0 + ci * chunk_b + k * 32, 0, 8u * rowB, p.layout);
            umma_bf16(tmem_base + kTmemS, da, db, idesc_s, kk > 0 ? 1u : 0u);
          }
        umma_commit(acc());
      } else {
        ++ev;
      }
      __syncwarp();
    };
    auto issue_pv = [&](int h) {
      if (elect_one_sync()) {
        const uint32_t v0 = ra_addr + 2u * TB + static_cast<uint32_t>(h) * tile_b, p0 = ra_addr + 3u * TB;
        for (int k = 0; k < kT / 16; ++k) {
          const uint64_t da = umma_desc(p0 + (k >> 2) * (kT * 128) + (k & 3) * 32, 0, 1024, 2);
          const uint64_t db = umma_desc(v0 + k * 16 * rowB, chunk_b, 8u * rowB, p.layout);
          umma_bf16(tmem_base, da, db, idesc_o, k > 0 ? 1u : 0u);
        }
        umma_commit(acc());
      } else {
        ++ev;
      }
      __syncwarp();
    };
    auto gemm = [&](int a_mode, int n, int k32, uint32_t tcol) {
      uint64_t* ab = &acc_bar[ev & 1u];
      ++ev;
      mma_pass(ra_addr, ring_addr, a_mode, n, k32, tmem_base + tcol, afull, aempty, p.a_slots, a_par, bfull, bempty,
               p.stages, r, ab);
    };
    for (int b = cluster_id; b < p.B; b += n_clusters) {
      sp();
      sp();
      for (int l = 0; l < p.layers; ++l) {
        if (NQP == 1) {
          gemm(0, 3 * DC, k32_d, 0);
        } else {
          gemm(0, 256, k32_d, 0);
          gemm(1, 128, k32_d, 256);
        }
        // attention: S(h+1) is issued right behind PV(h), so it runs while the compute warps drain O(h)
        mbar_wait(op_bar, (opn++) & 1u);      // Q, K, V tiles written
        tc_fence_after();
        issue_s(0);
        for (int h = 0; h < p.hpc; ++h) {
          mbar_wait(op_bar, (opn++) & 1u);    // P(h) written, S(h) consumed
          tc_fence_after();
          issue_pv(h);
          if (h + 1 < p.hpc) issue_s(h + 1);
        }
        sp();                                 // #A
        gemm(0, DC, k32_d, 0);                // out-proj
        sp();                                 // #B
        sp();                                 // #C
        for (int u = 0; u < NUP; ++u) gemm(u == 0 ? 0 : 1, 256, k32_d, static_cast<uint32_t>(u) * 256u);
        sp();                                 // #D
        gemm(2, DC, 4 * k32_d, 0);            // down-proj, A streamed
        sp();                                 // #E
        sp();                                 // #F
      }
    }
    if (pend) cluster_wait_acquire();
  } else if (warp == 2) {
    // ===================================================== activation (A) producer
    uint32_t e_par = 0u;
    auto cb = [&]() {
      cluster_arrive_release();
      cluster_wait_acquire();
    };
    auto load_resident = [&](int buf, int b) {
      if (elect_one_sync()) {
        fence_proxy_async_all();
        const int row = (buf * p.B + b) * kT;
        for (int j = 0; j < kch_d; ++j) {
          mbar_arrive_expect_tx(&afull[j], kAChunk);
          tma_load_2d(RA + j * kAChunk, &maps.abuf, &afull[j], j * 64, row);
        }
      }
      __syncwarp();
    };
    for (int b = cluster_id; b < p.B; b += n_clusters) {
      cb();
      cb();
      load_resident(0, b);                    // LN1(x) of layer 0
      for (int l = 0; l < p.layers; ++l) {
        cb();                                 // #A
        load_resident(1, b);                  // O
        cb();                                 // #B
        cb();                                 // #C
        load_resident(2, b);                  // LN2(x)
        cb();                                 // #D: hidden published -> stream it through the operand slots
        for (int ch = 0; ch < 4 * kch_d; ++ch) {
          const int slot = ch % p.a_slots;
          mbar_wait(&aempty[slot], ((e_par >> slot) & 1u) ^ 1u);
          e_par ^= 1u << slot;
          if (elect_one_sync()) {
            if (ch == 0) fence_proxy_async_all();
            mbar_arrive_expect_tx(&afull[slot], kAChunk);
            tma_load_2d(RA + slot * kAChunk, &maps.hbuf, &afull[slot], ch * 64, b * kT);
          }
          __syncwarp();
        }
        cb();                                 // #E
        cb();                                 // #F
        if (l + 1 < p.layers) load_resident(0, b);
      }
    }
  } else if (warp == 3) {
    const int n_img = (p.B - cluster_id + n_clusters - 1) / n_clusters;
    for (int i = 0; i < n_img * cpi; ++i) {
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  } else {
    // ===================================================== compute warps: thread (t, hh)
    const int q = warp & 3;                 // TMEM lane quarter
    const int hh = (warp - 4) >> 2;         // column half
    const int t = q * 32 + lane;            // token row
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int col0 = rank * DC + hh * NC;   // first residual-stream column of this thread
    const int dk = p.dk, rowB = p.rowB, cw = p.cw;
    const int swz = p.layout == 2 ? (t & 7) : (p.layout == 4 ? ((t >> 1) & 3) : ((t >> 2) & 1));
    const int tile_b = kT * dk * 2;
    uint32_t ev = 0u;
    float xv[NC];
    auto acc_wait = [&]() {
      mbar_wait(&acc_bar[ev & 1u], (ev >> 1) & 1u);
      ++ev;
    };
    auto cb = [&]() {
      cluster_arrive_release();
      cluster_wait_acquire();
    };
    // LayerNorm of the cluster-distributed rows: partial sums to every CTA (DSMEM), normalise the own slice.
    // dst_bf16 != null: bf16 operand slice for the next GEMM (published by the second barrier); else fp32 output.
    auto ln_step = [&](const float* gamma, const float* beta, float eps, __nv_bfloat16* dst_bf16, float* dst_f32) {
      float s = 0.f, sq = 0.f;
#pragma unroll
      for (int i = 0; i < NC; ++i) {
        s += xv[i];
        sq = fmaf(xv[i], xv[i], sq);
      }
      const uint32_t laddr = smem_u32(&stats[(rank * 2 + hh) * kT + t]);
      for (int r = 0; r < C; ++r) st_cluster_v2f32(mapa_u32(laddr, static_cast<uint32_t>(r)), s, sq);
      cb();
      float S = 0.f, Q = 0.f;
      for (int j = 0; j < 2 * C; ++j) {
        const float2 v = stats[j * kT + t];
        S += v.x;
        Q += v.y;
      }
      const float inv_d = 1.0f / static_cast<float>(d);
      const float mean = S * inv_d;
      const float var = fmaxf(Q * inv_d - mean * mean, 0.f);
      const float rstd = rsqrtf(var + eps);
      const float4* g4 = reinterpret_cast<const float4*>(gamma + col0);
      const float4* b4 = reinterpret_cast<const float4*>(beta + col0);
#pragma unroll
      for (int i = 0; i < NC; i += 8) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float4 gg = __ldg(g4 + (i >> 2) + j), bb = __ldg(b4 + (i >> 2) + j);
          f[4 * j + 0] = (xv[i + 4 * j + 0] - mean) * rstd * gg.x + bb.x;
          f[4 * j + 1] = (xv[i + 4 * j + 1] - mean) * rstd * gg.y + bb.y;
          f[4 * j + 2] = (xv[i + 4 * j + 2] - mean) * rstd * gg.z + bb.z;
          f[4 * j + 3] = (xv[i + 4 * j + 3] - mean) * rstd * gg.w + bb.w;
        }
        if (dst_bf16 != nullptr) {
          *reinterpret_cast<bf16x8*>(dst_bf16 + i) = pack8(f);
        } else {
          *reinterpret_cast<float4*>(dst_f32 + i) = make_float4(f[0], f[1], f[2], f[3]);
          *reinterpret_cast<float4*>(dst_f32 + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
        }
      }
      fence_proxy_async_all();
      cb();
    };
    // x += accumulator + bias  (out-proj / down-proj epilogue; accumulator columns [0, DC))
    auto residual_epilogue = [&](const float* bias) {
      acc_wait();
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < NC; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + lane_addr + static_cast<uint32_t>(hh * NC + c0), v);
        const float4* b4 = reinterpret_cast<const float4*>(bias + col0 + c0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 bb = __ldg(b4 + i);
          xv[c0 + 4 * i + 0] += __uint_as_float(v[4 * i + 0]) + bb.x;
          xv[c0 + 4 * i + 1] += __uint_as_float(v[4 * i + 1]) + bb.y;
          xv[c0 + 4 * i + 2] += __uint_as_float(v[4 * i + 2]) + bb.z;
          xv[c0 + 4 * i + 3] += __uint_as_float(v[4 * i + 3]) + bb.w;
        }
      }
      tc_fence_before();
    };

    for (int b = cluster_id; b < p.B; b += n_clusters) {
      const size_t row = static_cast<size_t>(b) * kT + t;
      {
        const float4* xin = reinterpret_cast<const float4*>(p.x_in + row * d + col0);
#pragma unroll
        for (int i = 0; i < NC / 4; ++i) {
          const float4 v = xin[i];
          xv[4 * i + 0] = v.x;
          xv[4 * i + 1] = v.y;
          xv[4 * i + 2] = v.z;
          xv[4 * i + 3] = v.w;
        }
      }
      __nv_bfloat16* a_ln1 = p.abuf + (static_cast<size_t>(0) * p.B * kT + row) * d + col0;
      __nv_bfloat16* a_o = p.abuf + (static_cast<size_t>(1) * p.B * kT + row) * d + rank * DC;
      __nv_bfloat16* a_ln2 = p.abuf + (static_cast<size_t>(2) * p.B * kT + row) * d + col0;
      ln_step(p.ln1g, p.ln1b, p.eps1, a_ln1, nullptr);

      for (int l = 0; l < p.layers; ++l) {
        // ---------------- QKV accumulator -> Q / K / V operand tiles (they overwrite the dead LN1(x) operand)
        for (int z = 0; z < NQP; ++z) acc_wait();
        tc_fence_after();
        {
          const float* bq = p.bqkv + l * 3 * d + rank * DC;
#pragma unroll 1
          for (int z = 0; z < NQP; ++z) {
            const int n = NQP == 1 ? 3 * DC : (z == 0 ? 256 : 128);
            const int g0 = z == 0 ? 0 : 256;                    // first q|k|v column of the pass (= its TMEM column)
#pragma unroll 1
            for (int c0 = hh * (n / 2); c0 < (hh + 1) * (n / 2); c0 += 32) {
              const int g = g0 + c0;                            // column within this CTA's [q | k | v], multiple of 32
              uint32_t v[32];
              tmem_ld32(tmem_base + lane_addr + static_cast<uint32_t>(g), v);
              const int part = g / DC, m0 = g - part * DC;
              const float4* b4 = reinterpret_cast<const float4*>(bq + part * d + m0);
              uint8_t* part_base = RA + part * TB;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b0 = __ldg(b4 + 2 * j), b1 = __ldg(b4 + 2 * j + 1);
                float f[8];
                f[0] = __uint_as_float(v[8 * j + 0]) + b0.x;
                f[1] = __uint_as_float(v[8 * j + 1]) + b0.y;
                f[2] = __uint_as_float(v[8 * j + 2]) + b0.z;
                f[3] = __uint_as_float(v[8 * j + 3]) + b0.w;
                f[4] = __uint_as_float(v[8 * j + 4]) + b1.x;
                f[5] = __uint_as_float(v[8 * j + 5]) + b1.y;
                f[6] = __uint_as_float(v[8 * j + 6]) + b1.z;
                f[7] = __uint_as_float(v[8 * j + 7]) + b1.w;
                const int m = m0 + 8 * j;
                const int head = m / dk, e = m - head * dk;
                const int ci = e / cw, ec = e - ci * cw;
                uint8_t* dst = part_base + head * tile_b + ci * (kT * rowB) + t * rowB + (((ec >> 3) ^ swz) << 4);
                *reinterpret_cast<bf16x8*>(dst) = pack8(f);
              }
            }
          }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(op_bar);

        // ---------------- attention over the local heads
        for (int h = 0; h < p.hpc; ++h) {
          acc_wait();                                  // S(h)
          tc_fence_after();
          float sum = 0.f;
          {
            uint32_t v0[32], v1[32];
            tmem_ld32(tmem_base + lane_addr + kTmemS + static_cast<uint32_t>(hh * 64), v0);
            tmem_ld32(tmem_base + lane_addr + kTmemS + static_cast<uint32_t>(hh * 64 + 32), v1);
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaxf(__uint_as_float(v0[i]), __uint_as_float(v1[i])));
            smax[hh * kT + t] = mx;
            named_bar_sync(1, 32 * kComputeWarps);
            mx = fmaxf(mx, smax[(hh ^ 1) * kT + t]);
            const float mxs = mx * p.scale_log2e;
            uint8_t* prow = RA + 3 * TB + hh * (kT * 128) + t * 128;     // P chunk hh (64 keys), row = query
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t raw = j < 4 ? v0[8 * j + i] : v1[8 * (j - 4) + i];
                const float e = exp2f(fmaf(__uint_as_float(raw), p.scale_log2e, -mxs));
                f[i] = __bfloat162float(__float2bfloat16_rn(e));
                sum += f[i];                               // normalise by what the MMA will see
              }
              *reinterpret_cast<bf16x8*>(prow + ((j ^ (t & 7)) << 4)) = pack8(f);
            }
            ssum[hh * kT + t] = sum;
          }
          fence_proxy_async();
          tc_fence_before();
          named_bar_sync(1, 32 * kComputeWarps);          // partner's row sum visible; all S reads retired
          if (lane == 0) mbar_arrive(op_bar);
          acc_wait();                                  // O(h)
          tc_fence_after();
          const float inv = 1.0f / (ssum[t] + ssum[kT + t]);
          __nv_bfloat16* orow = a_o + h * dk;
          if (dk >= 64) {
            for (int c0 = 0; c0 < dk / 2; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(tmem_base + lane_addr + static_cast<uint32_t>(hh * (dk / 2) + c0), v);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * j + i]) * inv;
                *reinterpret_cast<bf16x8*>(orow + hh * (dk / 2) + c0 + 8 * j) = pack8(f);
              }
            }
          } else {
            uint32_t v[32];
            tmem_ld32(tmem_base + lane_addr, v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (8 * j >= hh * (dk / 2) && 8 * j < (hh + 1) * (dk / 2)) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[8 * j + i]) * inv;
                *reinterpret_cast<bf16x8*>(orow + 8 * j) = pack8(f);
              }
            }
          }
          tc_fence_before();
        }
        fence_proxy_async_all();
        cb();                                                    // #A: O of every head published

        // ---------------- out-proj epilogue + LN2
        residual_epilogue(p.bo + l * d);
        ln_step(p.ln2g + l * d, p.ln2b + l * d, p.eps2, a_ln2, nullptr);       // #B, #C

        // ---------------- MLP up: +bias, erf-GELU -> hidden slice
        {
          __nv_bfloat16* hrow = p.hbuf + row * (4 * d) + rank * 4 * DC + hh * 128;
          const float* b1 = p.b1 + l * 4 * d + rank * 4 * DC + hh * 128;
#pragma unroll 1
          for (int u = 0; u < NUP; ++u) {
            acc_wait();
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(tmem_base + lane_addr + static_cast<uint32_t>(u * 256 + hh * 128 + c0), v);
              const float4* b4 = reinterpret_cast<const float4*>(b1 + u * 256 + c0);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b0 = __ldg(b4 + 2 * j), bb1 = __ldg(b4 + 2 * j + 1);
                float f[8];
                f[0] = gelu_fast(__uint_as_float(v[8 * j + 0]) + b0.x);
                f[1] = gelu_fast(__uint_as_float(v[8 * j + 1]) + b0.y);
                f[2] = gelu_fast(__uint_as_float(v[8 * j + 2]) + b0.z);
                f[3] = gelu_fast(__uint_as_float(v[8 * j + 3]) + b0.w);
                f[4] = gelu_fast(__uint_as_float(v[8 * j + 4]) + bb1.x);
                f[5] = gelu_fast(__uint_as_float(v[8 * j + 5]) + bb1.y);
                f[6] = gelu_fast(__uint_as_float(v[8 * j + 6]) + bb1.z);
                f[7] = gelu_fast(__uint_as_float(v[8 * j + 7]) + bb1.w);
                *reinterpret_cast<bf16x8*>(hrow + u * 256 + c0 + 8 * j) = pack8(f);
              }
            }
          }
          tc_fence_before();
        }
        fence_proxy_async_all();
        cb();                                                    // #D: hidden published

        // ---------------- down-proj epilogue + next LN1 / ln_f
        residual_epilogue(p.b2 + l * d);
        if (p.dbg != nullptr) {
          float* dp = p.dbg + ((static_cast<size_t>(l) * p.B * kT) + row) * d + col0;
#pragma unroll
          for (int i = 0; i < NC; ++i) dp[i] = xv[i];
        }
        if (l + 1 < p.layers)
          ln_step(p.ln1g + (l + 1) * d, p.ln1b + (l + 1) * d, p.eps1, a_ln1, nullptr);     // #E, #F
        else
          ln_step(p.lnfg, p.lnfb, p.epsf, nullptr, p.x_out + row * d + col0);              // #E, #F
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();          // no CTA leaves while a peer may still address its shared memory
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ host side
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

int encode2d(CUtensorMap* m, const void* base, long long cols, long long rows, int box_cols, int box_rows,
             CUtensorMapSwizzle swz, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return CFT_E_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, str, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s: %lld x %lld, box %d x %d) failed (CUresult %d)", what, rows, cols, box_rows,
              box_cols, (int)r);
    return CFT_E_CUDA;
  }
  return CFT_OK;
}

// Cluster size for (B, d): DC = d / C must be 64 or 128 and C must divide the head count.  Prefer the widest split
// whose B clusters still fit on the GPU at once (more SMs per image), else the narrowest.
int pick_cluster(int B, int d, int heads, int forced) {
  int best = 0;
  for (int c = 8; c >= 1; c >>= 1) {
    if (heads % c || d % c) continue;
    const int dc = d / c;
    if (dc != 64 && dc != 128) continue;
    if (forced) {
      if (c == forced) return c;
      continue;
    }
    if (!best) best = c;                               // widest legal split
    if (static_cast<long long>(B) * c <= sm_count() - 16) return c;
    best = c;                                          // remember the narrowest seen so far
  }
  return forced ? 0 : best;
}

struct Plan {
  int C, DC, ra_bytes, a_slots, stages, smem;
};
bool make_plan(int B, int d, int heads, int forced_c, Plan* pl) {
  if (heads <= 0 || d % heads) return false;
  const int dk = d / heads;
  if (dk != 16 && dk != 32 && dk != 64 && dk != 128) return false;
  if (d % 64 || d > 512) return false;
  const int C = pick_cluster(B, d, heads, forced_c);
  if (!C) return false;
  const int DC = d / C;
  int ra = kT * d * 2;
  const int tiles = 3 * kT * DC * 2 + 2 * kT * 128;
  if (tiles > ra) ra = tiles;
  ra = (ra + kAChunk - 1) / kAChunk * kAChunk;
  const int misc = 512 + 2048 + C * 2 * kT * 8;
  const int stages_max = (kSmemMax - 1024 - ra - misc) / kStageBytes;
  if (ra / kAChunk > kMaxASlots || ra / kAChunk < d / 64 || stages_max < 3) return false;
  pl->C = C;
  pl->DC = DC;
  pl->ra_bytes = ra;
  pl->a_slots = ra / kAChunk;
  pl->stages = stages_max > kMaxStages ? kMaxStages : stages_max;
  pl->smem = 1024 + ra + pl->stages * kStageBytes + misc;
  return true;
}

bool g_attr_set = false;
const int g_force_c = getenv("CFT_BLOCK_CLUSTER") ? atoi(getenv("CFT_BLOCK_CLUSTER")) : 0;

}  // namespace

using namespace cft;

extern "C" long long cft_gpt_block_workspace_bytes(int B, int d) {
  if (B <= 0 || d <= 0) return 0;
  return static_cast<long long>(B) * kT * d * 2 * (3 + 4);
}

extern "C" int cft_gpt_block_supported(int B, int d, int heads, int tokens) {
  Plan pl;
  return (tokens == kT && B > 0 && make_plan(B, d, heads, g_force_c, &pl)) ? 1 : 0;
}

extern "C" int cft_gpt_block(const cft_gpt_block_args* a, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CFT_REQUIRE(a != nullptr, "cft_gpt_block: null args");
  CFT_REQUIRE(a->wqkv && a->bqkv && a->wo && a->bo && a->w1 && a->b1 && a->w2 && a->b2 && a->ln1_g && a->ln1_b &&
                  a->ln2_g && a->ln2_b && a->lnf_g && a->lnf_b && a->x_in && a->x_out && a->workspace,
              "cft_gpt_block: null pointer");
  CFT_REQUIRE(a->B > 0 && a->layers > 0 && a->tokens == kT, "cft_gpt_block: need B > 0, layers > 0, 128 tokens per image");
  Plan pl;
  if (!make_plan(a->B, a->d, a->heads, a->cluster > 0 ? a->cluster : g_force_c, &pl)) {
    set_error("cft_gpt_block: shape outside the fused kernel (d %d heads %d): use the per-op path", a->d, a->heads);
    return CFT_E_UNSUPPORTED;
  }
  CFT_REQUIRE(a->workspace_bytes >= cft_gpt_block_workspace_bytes(a->B, a->d), "cft_gpt_block: workspace too small");
  CFT_REQUIRE(reinterpret_cast<uintptr_t>(a->workspace) % 128 == 0 && reinterpret_cast<uintptr_t>(a->wqkv) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(a->wo) % 16 == 0 && reinterpret_cast<uintptr_t>(a->w1) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(a->w2) % 16 == 0 && reinterpret_cast<uintptr_t>(a->x_in) % 16 == 0 &&
                  reinterpret_cast<uintptr_t>(a->x_out) % 16 == 0,
              "cft_gpt_block: misaligned pointer");
  const int d = a->d, L = a->layers, B = a->B;
  BlockParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.d = d; p.heads = a->heads; p.dk = d / a->heads; p.layers = L; p.C = pl.C; p.hpc = a->heads / pl.C;
  p.stages = pl.stages; p.a_slots = pl.a_slots; p.ra_bytes = pl.ra_bytes;
  p.cw = p.dk < 64 ? p.dk : 64;
  p.nch = p.dk / p.cw;
  p.rowB = p.cw * 2;
  p.layout = p.cw == 64 ? 2 : (p.cw == 32 ? 4 : 6);
  p.scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(p.dk));
  p.eps1 = a->eps1; p.eps2 = a->eps2; p.epsf = a->epsf;
  p.bqkv = a->bqkv; p.bo = a->bo; p.b1 = a->b1; p.b2 = a->b2;
  p.ln1g = a->ln1_g; p.ln1b = a->ln1_b; p.ln2g = a->ln2_g; p.ln2b = a->ln2_b; p.lnfg = a->lnf_g; p.lnfb = a->lnf_b;
  p.x_in = a->x_in; p.x_out = a->x_out;
  p.abuf = reinterpret_cast<__nv_bfloat16*>(a->workspace);
  p.hbuf = p.abuf + static_cast<size_t>(3) * B * kT * d;
  p.dbg = a->debug_x;

  BlockMaps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  if ((rc = encode2d(&maps.wqkv, a->wqkv, d, static_cast<long long>(L) * 3 * d, 32, pl.DC, CU_TENSOR_MAP_SWIZZLE_64B, "wqkv"))) return rc;
  if ((rc = encode2d(&maps.wo, a->wo, d, static_cast<long long>(L) * d, 32, pl.DC, CU_TENSOR_MAP_SWIZZLE_64B, "wo"))) return rc;
  if ((rc = encode2d(&maps.w1, a->w1, d, static_cast<long long>(L) * 4 * d, 32, 256, CU_TENSOR_MAP_SWIZZLE_64B, "w1"))) return rc;
  if ((rc = encode2d(&maps.w2, a->w2, 4 * d, static_cast<long long>(L) * d, 32, pl.DC, CU_TENSOR_MAP_SWIZZLE_64B, "w2"))) return rc;
  if ((rc = encode2d(&maps.abuf, p.abuf, d, static_cast<long long>(3) * B * kT, 64, kT, CU_TENSOR_MAP_SWIZZLE_128B, "abuf"))) return rc;
  if ((rc = encode2d(&maps.hbuf, p.hbuf, 4 * d, static_cast<long long>(B) * kT, 64, kT, CU_TENSOR_MAP_SWIZZLE_128B, "hbuf"))) return rc;

  if (!g_attr_set) {
    rc = check_cuda(cudaFuncSetAttribute(cft_gpt_block_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax),
                    "cudaFuncSetAttribute(gpt_block<64>)");
    if (rc) return rc;
    rc = check_cuda(cudaFuncSetAttribute(cft_gpt_block_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax),
                    "cudaFuncSetAttribute(gpt_block<128>)");
    if (rc) return rc;
    g_attr_set = true;
  }
  int clusters = B;
  const int max_clusters = sm_count() / pl.C;
  if (clusters > max_clusters) clusters = max_clusters;
  LaunchScope ls(CFT_K_GPT_BLOCK, stream);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(clusters * pl.C);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pl.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = pl.DC == 64 ? cudaLaunchKernelEx(&cfg, cft_gpt_block_kernel<64>, maps, p)
                              : cudaLaunchKernelEx(&cfg, cft_gpt_block_kernel<128>, maps, p);
  if (e != cudaSuccess) {
    ls.finish("cft_gpt_block launch");
    return check_cuda(e, "cudaLaunchKernelEx(gpt_block)");
  }
  return ls.finish("cft_gpt_block launch");
}
