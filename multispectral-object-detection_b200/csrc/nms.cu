// Batched non-maximum suppression on Detect's output z [B, rows, 5 + nc] (fp32), one CTA per image:
//
//   1. filter   obj > conf_thres, conf = cls * obj (best class, or every class with multi_label) > conf_thres,
//               optional class filter  ->  64-bit keys {~bits(conf), candidate id} appended to the image's key list
//   2. sort     bitonic network (all-ascending form, so the virtual +inf padding never moves) in shared memory
//               (<= 26 k candidates; 640x640 has 25200 rows) or in the L2-resident workspace beyond that:
//               conf descending, candidate id ascending == the stable descending sort of torchvision.ops.nms
//   3. greedy   1024 candidates per chunk, one per thread: test against the kept list, then resolve the chunk 32
//               candidates per round (the 32x32 IoU bit matrix is computed by all 32 warps, one IoU per thread; the
//               serial scan over its rows is replayed by every warp), the newly kept boxes go through smem to the
//               threads that still hold live candidates; stops at max_det kept boxes (the reference truncates afterwards: same result,
//               a greedy decision depends on earlier kept boxes only)
//
// Integer/index work is exact; the fp32 arithmetic uses the reference's operation order with explicit round-to-nearest
// intrinsics (no FMA contraction), so the output equals utils/general.py:455-544 + torchvision's CPU kernel bit for
// bit (tests/test_nms_gpu.py).  Replaces reference utils/general.py:455-544 (non_max_suppression), :299-306
// (xywh2xyxy) and torchvision.ops.nms (:527).
#include "cft_common.cuh"

namespace {
using namespace cft;

constexpr int kNmsThreads = 1024;
constexpr int kNmsWarps = kNmsThreads / 32;
static_assert(kNmsWarps == 32, "the round loop maps warp w to lane w of every warp (s_alive[lane], matrix row = warp)");
constexpr int kMaxDetCap = 1024;
constexpr int kMaxNms = 30000;             // utils/general.py:466
constexpr float kMaxWh = 4096.f;           // utils/general.py:464

struct NmsArgs {
  const float* pred;                // [B][rows][no]
  int rows, no, nc;
  float conf_thres, iou_thres;
  int max_det, multi_label, agnostic;
  unsigned long long cls_mask[4];   // allowed classes (bit c), all ones = no filter
  unsigned long long* keys_ws;      // [B][cap]
  long long cap;                    // rows * (multi_label ? nc : 1)
  int smem_keys;                    // key slots available in shared memory
  float* out;                       // [B][max_det][6]
  int* counts;                      // [B]
};

struct Cand {
  float x1, y1, x2, y2;             // class-offset box (what torchvision.ops.nms sees)
  float area;
};

__device__ __forceinline__ bool iou_gt(const Cand& a, float bx1, float by1, float bx2, float by2, float barea, float thr) {
  // a = the kept (earlier) box i, b = the later box j:  inter / (area_i + area_j - inter) > thr
  const float w = fmaxf(0.f, __fsub_rn(fminf(a.x2, bx2), fmaxf(a.x1, bx1)));
  const float h = fmaxf(0.f, __fsub_rn(fminf(a.y2, by2), fmaxf(a.y1, by1)));
  const float inter = __fmul_rn(w, h);
  // disjoint boxes (the common case): 0 / x is 0, -0 or NaN, never > thr >= 0 -- skip the IEEE division (its operand
  // check sends a zero numerator down the slow path)
  if (inter == 0.f && thr >= 0.f) return false;
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, barea), inter));
  return ovr > thr;
}

__global__ void __launch_bounds__(kNmsThreads, 1) nms_kernel(const __grid_constant__ NmsArgs a) {
  pdl_prologue();
  extern __shared__ unsigned long long nms_smem[];
  __shared__ int s_n, s_kept;
  __shared__ unsigned s_alive[kNmsWarps], s_sup[32];
  __shared__ float s_batch[5][32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pred = a.pred + static_cast<long long>(b) * a.rows * a.no;
  unsigned long long* ws = a.keys_ws + static_cast<long long>(b) * a.cap;
  // smem layout: kept list (5 floats x max_det) first, keys after it
  float* kx1 = reinterpret_cast<float*>(nms_smem);
  float* ky1 = kx1 + a.max_det;
  float* kx2 = ky1 + a.max_det;
  float* ky2 = kx2 + a.max_det;
  float* kar = ky2 + a.max_det;
  unsigned long long* skeys = nms_smem + (5 * a.max_det + 1) / 2;
  if (tid == 0) {
    s_n = 0;
    s_kept = 0;
  }
  __syncthreads();

  // ---- 1. filter -> keys (any order; the sort restores candidate order through the id in the low word)
  const int ncand = a.multi_label ? a.nc : 1;
  for (int r = tid; r < a.rows; r += kNmsThreads) {
    const float* p = pred + static_cast<long long>(r) * a.no;
    const float obj = p[4];
    if (!(obj > a.conf_thres)) continue;
    if (a.multi_label) {
      for (int j = 0; j < a.nc; ++j) {
        const float c = __fmul_rn(p[5 + j], obj);
        if (c > a.conf_thres && ((a.cls_mask[j >> 6] >> (j & 63)) & 1ull)) {
          const int slot = atomicAdd(&s_n, 1);
          ws[slot] = (static_cast<unsigned long long>(~__float_as_uint(c)) << 32) | static_cast<unsigned>(r * ncand + j);
        }
      }
    } else {
      float best = __fmul_rn(p[5], obj);
      int bj = 0;
      for (int j = 1; j < a.nc; ++j) {
        const float c = __fmul_rn(p[5 + j], obj);
        if (c > best) {               // first maximum wins (torch.max / argmax)
          best = c;
          bj = j;
        }
      }
      if (best > a.conf_thres && ((a.cls_mask[bj >> 6] >> (bj & 63)) & 1ull)) {
        const int slot = atomicAdd(&s_n, 1);
        ws[slot] = (static_cast<unsigned long long>(~__float_as_uint(best)) << 32) | static_cast<unsigned>(r);
      }
    }
  }
  __syncthreads();
  int n = s_n;

  // ---- 2. sort ascending by key = (conf descending, candidate id ascending)
  unsigned long long* keys = ws;
  if (n <= a.smem_keys) {
    for (int i = tid; i < n; i += kNmsThreads) skeys[i] = ws[i];
    keys = skeys;
  }
  __syncthreads();
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  // pair t of a step: (i, l) with i < l; all strides are powers of two -> shifts and masks only.  Four pairs per thread
  // are loaded before any is stored (the pairs of one step are disjoint, which the compiler cannot know).
  const int half = n2 >> 1;
  auto step = [&](auto pair_of) {
    for (int t0 = tid; t0 < half; t0 += 4 * kNmsThreads) {
      int pi[4], pl[4];
      unsigned long long va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int t = t0 + u * kNmsThreads;
        pi[u] = -1;
        if (t < half) {
          int i, l;
          pair_of(t, i, l);
          if (l < n) {
            pi[u] = i;
            pl[u] = l;
            va[u] = keys[i];
            vb[u] = keys[l];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (pi[u] >= 0 && va[u] > vb[u]) {
          keys[pi[u]] = vb[u];
          keys[pl[u]] = va[u];
        }
    }
    __syncthreads();
  };
  for (int k = 2, lk = 1; k <= n2; k <<= 1, ++lk) {
    const int hk = k >> 1;
    step([&](int t, int& i, int& l) {                           // flip step: i <-> block end - offset
      const int blk = t >> (lk - 1), off = t & (hk - 1);
      i = (blk << lk) + off;
      l = (blk << lk) + k - 1 - off;
    });
    for (int j = hk >> 1; j >= 1; j >>= 1)                      // half cleaners
      step([&](int t, int& i, int& l) {
        i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        l = i + j;
      });
  }
  if (n > kMaxNms) n = kMaxNms;                                 // utils/general.py:521-522 (top max_nms by confidence)

  // ---- 3. greedy suppression, 1024 sorted candidates per chunk
  const float thr = a.iou_thres;
  float* out = a.out + static_cast<long long>(b) * a.max_det * 6;
  for (int base = 0; base < n; base += kNmsThreads) {
    if (s_kept >= a.max_det) break;                             // uniform: s_kept only changes between barriers
    const int pos = base + tid;
    bool alive = pos < n;
    Cand c{0.f, 0.f, 0.f, 0.f, 0.f};
    float ux1 = 0.f, uy1 = 0.f, ux2 = 0.f, uy2 = 0.f, conf = 0.f, clsf = 0.f;
    if (alive) {
      const unsigned long long key = keys[pos];
      const unsigned id = static_cast<unsigned>(key);
      conf = __uint_as_float(~static_cast<unsigned>(key >> 32));
      const int r = static_cast<int>(id / static_cast<unsigned>(ncand));
      int j = static_cast<int>(id - static_cast<unsigned>(r) * ncand);
      const float* p = pred + static_cast<long long>(r) * a.no;
      if (!a.multi_label) {                                     // best class again (cheap; keeps the key at 64 bits)
        const float obj = p[4];
        float best = __fmul_rn(p[5], obj);
        for (int q = 1; q < a.nc; ++q) {
          const float v = __fmul_rn(p[5 + q], obj);
          if (v > best) {
            best = v;
            j = q;
          }
        }
      }
      clsf = static_cast<float>(j);
      const float hw = __fmul_rn(p[2], 0.5f), hh = __fmul_rn(p[3], 0.5f);     // w / 2 is exact either way
      ux1 = __fsub_rn(p[0], hw);
      uy1 = __fsub_rn(p[1], hh);
      ux2 = __fadd_rn(p[0], hw);
      uy2 = __fadd_rn(p[1], hh);
      const float off = a.agnostic ? 0.f : __fmul_rn(clsf, kMaxWh);            // :525
      c.x1 = __fadd_rn(ux1, off);
      c.y1 = __fadd_rn(uy1, off);
      c.x2 = __fadd_rn(ux2, off);
      c.y2 = __fadd_rn(uy2, off);
      c.area = __fmul_rn(__fsub_rn(c.x2, c.x1), __fsub_rn(c.y2, c.y1));
      const int kept0 = s_kept;
      for (int k = 0; k < kept0 && alive; ++k) {
        const Cand kb{kx1[k], ky1[k], kx2[k], ky2[k], kar[k]};
        if (iou_gt(kb, c.x1, c.y1, c.x2, c.y2, c.area, thr)) alive = false;
      }
    }
    // resolve the chunk: one warp's 32 candidates per round (the first warp that still has live ones).  All 32 warps
    // take part: warp w computes row w of the batch's 32x32 suppression matrix (one IoU per thread), every warp then
    // replays the short serial scan over the 32 row masks, so no result has to be broadcast.
    for (;;) {
      const unsigned m = __ballot_sync(0xffffffffu, alive);
      if (lane == 0) s_alive[warp] = m;
      __syncthreads();
      const int kept_before = s_kept;
      if (kept_before >= a.max_det) break;
      const unsigned wm = s_alive[lane];                        // lane l looks at warp l (kNmsWarps == 32)
      const unsigned nz = __ballot_sync(0xffffffffu, wm != 0u);
      if (nz == 0u) break;
      const int fw = __ffs(nz) - 1;
      const unsigned fm = __shfl_sync(0xffffffffu, wm, fw);     // live candidates of the batch
      if (warp == fw) {
        s_batch[0][lane] = c.x1; s_batch[1][lane] = c.y1; s_batch[2][lane] = c.x2; s_batch[3][lane] = c.y2;
        s_batch[4][lane] = c.area;
      }
      __syncthreads();
      {
        const Cand r{s_batch[0][warp], s_batch[1][warp], s_batch[2][warp], s_batch[3][warp], s_batch[4][warp]};
        const bool sup = lane > warp &&
                         iou_gt(r, s_batch[0][lane], s_batch[1][lane], s_batch[2][lane], s_batch[3][lane], s_batch[4][lane], thr);
        const unsigned row = __ballot_sync(0xffffffffu, sup);
        if (lane == 0) s_sup[warp] = row;
      }
      __syncthreads();
      unsigned remaining = fm, keptmask = 0;
      int room = a.max_det - kept_before;
      while (remaining && room > 0) {
        const int l = __ffs(remaining) - 1;
        keptmask |= 1u << l;
        --room;
        remaining &= ~(1u << l);
        remaining &= ~s_sup[l];
      }
      const int kept_now = kept_before + __popc(keptmask);
      if (warp == fw) {
        if ((keptmask >> lane) & 1u) {
          const int k = kept_before + __popc(keptmask & ((1u << lane) - 1u));
          kx1[k] = c.x1; ky1[k] = c.y1; kx2[k] = c.x2; ky2[k] = c.y2; kar[k] = c.area;
          float* o = out + k * 6;
          o[0] = ux1; o[1] = uy1; o[2] = ux2; o[3] = uy2; o[4] = conf; o[5] = clsf;
        }
        alive = false;                                          // every candidate of this warp is decided
        if (lane == 0) s_kept = kept_now;
      }
      __syncthreads();
      if (alive) {
        for (int k = kept_before; k < kept_now && alive; ++k) {
          const Cand kb{kx1[k], ky1[k], kx2[k], ky2[k], kar[k]};
          if (iou_gt(kb, c.x1, c.y1, c.x2, c.y2, c.area, thr)) alive = false;
        }
      }
    }
    __syncthreads();
  }
  if (tid == 0) a.counts[b] = s_kept;
}

bool g_nms_attr = false;
}  // namespace

using namespace cft;

extern "C" long long cft_nms_workspace_bytes(int B, int rows, int nc, int multi_label) {
  if (B <= 0 || rows <= 0 || nc <= 0) return 0;
  const long long cap = static_cast<long long>(rows) * (multi_label && nc > 1 ? nc : 1);
  return static_cast<long long>(B) * cap * 8;
}

extern "C" int cft_nms(const float* pred, int B, int rows, int no, float conf_thres, float iou_thres, int max_det,
                       int multi_label, int agnostic, const int* classes, int n_classes, void* workspace,
                       long long workspace_bytes, float* out, int* counts, void* stream_v) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  CFT_REQUIRE(pred && out && counts && workspace, "cft_nms: null pointer");
  CFT_REQUIRE(B > 0 && rows > 0 && no > 5, "cft_nms: bad shape (B %d rows %d no %d)", B, rows, no);
  const int nc = no - 5;
  CFT_REQUIRE(nc <= 256, "cft_nms: at most 256 classes");
  CFT_REQUIRE(max_det > 0 && max_det <= kMaxDetCap, "cft_nms: max_det must be in [1, %d]", kMaxDetCap);
  CFT_REQUIRE(n_classes >= 0 && (n_classes == 0 || classes), "cft_nms: bad class filter");
  NmsArgs a;
  a.pred = pred;
  a.rows = rows;
  a.no = no;
  a.nc = nc;
  a.conf_thres = conf_thres;
  a.iou_thres = iou_thres;
  a.max_det = max_det;
  a.multi_label = (multi_label && nc > 1) ? 1 : 0;          // utils/general.py:471
  a.agnostic = agnostic ? 1 : 0;
  for (int i = 0; i < 4; ++i) a.cls_mask[i] = n_classes ? 0ull : ~0ull;
  for (int i = 0; i < n_classes; ++i)
    if (classes[i] >= 0 && classes[i] < 256) a.cls_mask[classes[i] >> 6] |= 1ull << (classes[i] & 63);
  a.cap = static_cast<long long>(rows) * (a.multi_label ? nc : 1);
  CFT_REQUIRE(a.cap < (1LL << 31), "cft_nms: too many candidates per image");
  CFT_REQUIRE(workspace_bytes >= static_cast<long long>(B) * a.cap * 8, "cft_nms: workspace too small (%lld < %lld bytes)",
              workspace_bytes, static_cast<long long>(B) * a.cap * 8);
  CFT_REQUIRE(reinterpret_cast<uintptr_t>(workspace) % 8 == 0, "cft_nms: workspace must be 8-byte aligned");
  a.keys_ws = static_cast<unsigned long long*>(workspace);
  a.out = out;
  a.counts = counts;
  const int kept_slots = (5 * max_det + 1) / 2;             // 8-byte units
  const long long max_keys = (220 * 1024) / 8 - kept_slots;
  a.smem_keys = static_cast<int>(a.cap < max_keys ? a.cap : max_keys);
  const size_t smem = static_cast<size_t>(kept_slots + a.smem_keys) * 8;
  if (!g_nms_attr) {
    int rc = check_cuda(cudaFuncSetAttribute(nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024),
                        "cudaFuncSetAttribute(nms)");
    if (rc) return rc;
    g_nms_attr = true;
  }
  LaunchScope ls(CFT_K_NMS, stream);
  cft::launch(nms_kernel, dim3(B), dim3(kNmsThreads), smem, stream, a);
  return ls.finish("cft_nms launch");
}
