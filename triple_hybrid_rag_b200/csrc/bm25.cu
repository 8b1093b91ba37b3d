// bm25.cu — K2: BM25 top-k over a CSR inverted index with per-term range skips (lexical channel).
//
// Stands where RAG2Retriever._lexical_search calls the rag2_lexical_search RPC
//   (src/voice_agent/rag2/retrieval.py:273-292, database/migrations/20260114_rag2_schema.sql:341-374):
//   same interface (keyword list in, top-`limit` by descending score out); the scoring formula is
//   BM25 as BASELINE.json's north_star asks.  Definition (oracle/bm25.py restates it on the CPU):
//     score(d) = fp32 sum, over the query's terms in order, of  idf[t] * impact(t, d)
//   with fp32 round-to-nearest multiply and add and no FMA contraction, so the result does not
//   depend on scheduling; docs with score > 0 are ranked by (score desc, id asc).
//
// Index (include/thr.h): postings {u32 doc, f32 impact} in term-major CSR order (doc ascending inside
// a term) and skip[t * n_blk + r] = first posting of term t whose doc lies in range r (blk_docs docs).
//
// Launch sequence of thr_bm25_topk: cost per query -> plan (cut queries into units = doc-range slices of
// about equal cost) -> order (heaviest first) -> bm25_span_kernel (one unit at a time per CTA, fetched
// dynamically; writes each unit's sorted top-k) -> bm25_merge_kernel (per query, merge its units' lists).
// The kernel is bound by instruction issue and latency, not by DRAM: a query touches ~13% of the docs, so
// what counts is the number of (term, span) visits and the instructions each visit costs.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kMaxTerms = 32;                    // lanes of the producer warp: one per query term
constexpr int kMaxBlkDocs = 2048;                // largest skip range
constexpr int kMaxSelB = 256;                    // largest k

struct Posting { uint32_t doc; float imp; };

// A work unit: one query restricted to the doc ranges [r0, r1).
struct Unit { int q; int r0; int r1; unsigned cost; };
constexpr int kMaxUnitsPerQuery = 64;    // small batches: one query can still spread over many SMs

struct Bm25Args {
  const int64_t* skip;    // [V * n_blk + 1]
  const Posting* post;
  const float* idf;
  int64_t n_docs;
  int n_blk, blk_docs, blk_shift, V;
  const int32_t* q_terms;
  const int32_t* q_off;
  const int32_t* order;   // work units, heaviest first
  const Unit* units;
  const int* total_units;
  int* work_counter;
  int B, k;
  uint64_t* part_keys;    // [max_units][k] sorted descending
  int32_t* part_cnt;      // [max_units]
  const uint16_t* tags;   // [n_docs] nullable: per-doc tag (collection id) for filtered queries
  const int32_t* want;    // [B] nullable: tag a query's docs must carry, < 0 = any
  thr_dev_status* status;
};

// Pull [p, p + bytes) into L2 (16-byte granules); no destination, no completion to wait for.
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Barrier among the first kNT threads of the CTA (named barrier kBar; kBar 0 with kNT = blockDim is __syncthreads).
template <int kNT, int kBar>
__device__ __forceinline__ void bar_group() {
  asm volatile("bar.sync %0, %1;" ::"n"(kBar), "n"(kNT) : "memory");
}

// Cooperative among kNT threads: keep the ksel largest of keys[0..n) in place, n > ksel, n <= kCap.
// Returns the ksel-th largest key.  hist/scal are shared scratch.
template <int kNT, int kCap, int kBar>
__device__ uint64_t block_compact_topk_t(uint64_t* keys, int n, int ksel, uint32_t* hist,
                                         unsigned long long* s_prefix, int* s_want, int* s_cnt, int tid) {
  if (tid == 0) { *s_prefix = 0ull; *s_want = ksel; }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = tid; i < 256; i += kNT) hist[i] = 0;
    bar_group<kNT, kBar>();
    const uint64_t prefix = *s_prefix;
    for (int i = tid; i < n; i += kNT) {
      const uint64_t key = keys[i];
      const bool match = pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8));
      if (match) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
    }
    bar_group<kNT, kBar>();
    // Find the digit that holds the want-th largest key (warp 0; common.cuh: radix_find_digit).
    if (tid < 32) {
      const int want = *s_want;
      int d, above;
      if (radix_find_digit(hist, want, tid, &d, &above)) {
        *s_want = want - above;
        *s_prefix = prefix | ((unsigned long long)d << shift);
      }
    }
    bar_group<kNT, kBar>();
  }
  const uint64_t T = *s_prefix;
  // survivors: read everything first, then rewrite the front
  constexpr int kPer = (kCap + kNT - 1) / kNT;
  uint64_t mine[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    int i = tid + j * kNT;
    mine[j] = i < n ? keys[i] : 0ull;
  }
  if (tid == 0) *s_cnt = 0;
  bar_group<kNT, kBar>();
#pragma unroll
  for (int j = 0; j < kPer; ++j)
    if (mine[j] >= T && mine[j] != 0ull) keys[atomicAdd(s_cnt, 1)] = mine[j];
  bar_group<kNT, kBar>();
  return T;
}
// ------------------------------------------------------------------------------------------------
// The CTA owns the accumulator.
//
// A span is kSpanDocs consecutive docs (a whole number of skip ranges) with one fp32 accumulator slot per doc
// in shared memory.  A term's postings inside a span are ONE contiguous piece of the term-major posting array
// (skip[t][first range] .. skip[t][last range + 1]), so the producer warp streams it into a ring of kChunk-posting
// stages with cp.async.bulk, one descriptor per chunk; all per-(term, span) bookkeeping lives in that one warp.
// The kSpanThreads consumer threads do nothing but: wait for a chunk, add its postings (<= 2 per thread, docs of
// one term are distinct: no atomics), release the stage.  Terms are separated by a named barrier (fp32 adds in
// query order: the sum is part of the definition); after the span's last chunk the accumulator is scanned
// with 16-byte loads: nonzero slots are zeroed and scores above tau are appended to the candidate list.
// An append that does not fit leaves its slot in place and raises a flag; the list is then compacted (which raises
// tau) and the scan repeats, so no threshold warm-up is needed for correctness.
//
// Shape (measured at 10M docs, batch 256; the THR_SPAN_* macros exist for such sweeps, scripts/build_variant.sh):
// time is (number of chunk visits) x (consumer warps) x (instructions per visit) at an IPC of ~1.7, so the span is
// as large as shared memory allows (fewer visits) and the common path of a visit is ~25 instructions; more
// warps hide more latency even though most of them have no posting in most chunks (a chunk holds a few hundred):
//   8 warps x 32k docs 3.36 ms | 16 x 32k 2.28 | 24 x 30k 2.18 | 2 CTAs x 8 warps x 16k 2.76 | 3 x 4 x 12k 3.15.
#ifndef THR_SPAN_DOCS
#define THR_SPAN_DOCS 30720
#endif
#ifndef THR_SPAN_WARPS
#define THR_SPAN_WARPS 24
#endif
#ifndef THR_SPAN_CHUNK
#define THR_SPAN_CHUNK 1536
#endif
#ifndef THR_SPAN_STAGES
#define THR_SPAN_STAGES 6
#endif
#ifndef THR_SPAN_CAP
#define THR_SPAN_CAP 2048
#endif
#ifndef THR_SPAN_CTAS
#define THR_SPAN_CTAS 1
#endif
constexpr int kSpanDocs = THR_SPAN_DOCS;          // a multiple of 2048 (every legal blk_docs divides it)
constexpr int kSpanWarps = THR_SPAN_WARPS;
constexpr int kSpanThreads = kSpanWarps * 32;     // consumers; warp kSpanWarps is the producer
constexpr int kChunk = THR_SPAN_CHUNK;            // postings per ring stage
constexpr int kStages = THR_SPAN_STAGES;
constexpr int kSpanCap = THR_SPAN_CAP;            // candidate slots
constexpr int kSpanCtas = THR_SPAN_CTAS;          // CTAs per SM
constexpr int kSpanBar = 1;                       // named barrier of the consumers
constexpr int kPerThread = kChunk / kSpanThreads; // postings per consumer thread and chunk
constexpr int kScanIters = kSpanDocs / (4 * kSpanThreads);
#ifndef THR_SPAN_ISSUE
#define THR_SPAN_ISSUE (THR_SPAN_STAGES / 2)
#endif
constexpr int kIssue = THR_SPAN_ISSUE;   // chunks the producer issues per round (half the ring)
constexpr int kScanBatch = kScanIters % 8 == 0 ? 8 : kScanIters % 6 == 0 ? 6 : kScanIters % 4 == 0 ? 4 : 1;
static_assert(kScanIters % kScanBatch == 0, "scan batches");
static_assert(kChunk % kSpanThreads == 0 && kChunk % 2 == 0 && kPerThread >= 1 && kPerThread <= 8, "chunk shape");
static_assert(kSpanDocs % kMaxBlkDocs == 0 && kSpanDocs % (4 * kSpanThreads) == 0, "span shape");
static_assert(kSpanThreads >= kMaxSelB / 2, "the final sort uses kMaxSelB / 2 threads");
static_assert(kSpanCap >= 2 * kMaxSelB + 256, "a compaction must free a useful part of the list");

enum : uint32_t { kFNewUnit = 1u, kFSync = 2u, kFEndSpan = 4u, kFEndUnit = 8u, kFExit = 16u };

constexpr size_t kSpanSmem = (size_t)kSpanDocs * 4 + (size_t)kSpanCap * 8 + (size_t)kStages * kChunk * 8 +
                             kStages * 16 + 2 * kStages * 8 + 256 * 4 + 64 + 256;
static_assert(kSpanCtas * (kSpanSmem + 1024) <= 233472, "kSpanCtas CTAs per SM");

__global__ void __launch_bounds__(kSpanThreads + 32, kSpanCtas) bm25_span_kernel(const Bm25Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* acc = (float*)gen;                                                   // [kSpanDocs]
  uint64_t* cand = (uint64_t*)(acc + kSpanDocs);                              // [kSpanCap]
  Posting* ring = (Posting*)(cand + kSpanCap);                               // [kStages][kChunk]
  uint4* desc = (uint4*)(ring + kStages * kChunk);                            // [kStages]
  uint64_t* full = (uint64_t*)(desc + kStages);                               // [kStages]
  uint64_t* empty = full + kStages;                                           // [kStages]
  uint32_t* hist = (uint32_t*)(empty + kStages);                              // 256
  unsigned long long* s_prefix = (unsigned long long*)(hist + 256);
  int* s_int = (int*)(s_prefix + 1);  // [0]=want [1]=cnt(compact) [2]=cand count [3]=overflow

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < kSpanDocs; i += kSpanThreads + 32) acc[i] = 0.f;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), kSpanWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_int[2] = 0;
    s_int[3] = 0;
  }
  __syncthreads();
  const uint32_t ring_u = smem_u32(ring), desc_u = smem_u32(desc), full_u = smem_u32(full), empty_u = smem_u32(empty);
  const int spr = kSpanDocs / a.blk_docs;   // ranges per span

  if (warp == kSpanWarps) {
    // ================= producer: lane t <-> query term t =================
    uint32_t s = 0, ph = 0;
    auto put = [&](uint32_t x, uint32_t y, uint32_t z, uint32_t w, const Posting* src, uint32_t bytes) {
      mbar_wait_relaxed(empty_u + s * 8u, ph ^ 1u, a.status, 470);
      if (lane == 0) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(desc_u + s * 16u), "r"(x), "r"(y), "r"(z), "r"(w)
                     : "memory");
        if (bytes) {
          mbar_arrive_expect_tx(full_u + s * 8u, bytes);
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              :
              : "r"(ring_u + s * (uint32_t)(kChunk * 8)), "l"(src), "r"(bytes), "r"(full_u + s * 8u)
              : "memory");
        } else {
          mbar_arrive(full_u + s * 8u);
        }
      }
      __syncwarp();
      if (++s == kStages) { s = 0; ph ^= 1u; }
    };
    for (;;) {
      int unit = -1;
      if (lane == 0) {
        const int w = atomicAdd(a.work_counter, 1);
        unit = w < *a.total_units ? a.order[w] : -1;
      }
      unit = __shfl_sync(0xffffffffu, unit, 0);
      if (unit < 0) break;
      const int q = a.units[unit].q, r0 = a.units[unit].r0, r1 = a.units[unit].r1;
      const int qlo = a.q_off[q];
      const int nt = min(a.q_off[q + 1] - qlo, kMaxTerms);   // the host rejects longer queries
      int term = -1;
      float wgt = 0.f;
      if (lane < nt) {
        term = a.q_terms[qlo + lane];
        if (term < 0 || term >= a.V) term = -1; else wgt = a.idf[term];
      }
      const int64_t* row = a.skip + (size_t)(term < 0 ? 0 : term) * a.n_blk;
      auto edge = [&](int i) -> int64_t {   // first posting of this lane's term at the start of span i of the unit
        return term >= 0 ? __ldg(row + min(r0 + i * spr, r1)) : 0;
      };
      const int nsp = (r1 - r0 + spr - 1) / spr;
      int64_t e0 = edge(0), e1 = edge(1), e2 = edge(2), e3 = edge(3);
      bool first = true;
      for (int sp = 0; sp < nsp; ++sp) {
        const int64_t e4 = edge(sp + 4);
        // the span after the next one -> L2 (the ring then pulls from L2, not DRAM); capped per term
        if (e3 > e2) {
          const int64_t b = e2 & ~(int64_t)1;
          const int64_t nby = min((e3 - b) * 8, (int64_t)32768);
          prefetch_l2_bulk(a.post + b, (uint32_t)((nby + 15) & ~(int64_t)15));
        }
        const int cnt = (int)(e1 - e0);
        const unsigned live = __ballot_sync(0xffffffffu, cnt > 0);
        if (live) {
          const uint32_t doc0 = (uint32_t)(r0 + sp * spr) << a.blk_shift;
          const int first_live = __ffs(live) - 1;
          // Chunks of this span, generated lane-parallel.  Term t (lane t) owns nch chunks: the first one ends at a
          // chunk boundary of the posting array's even-aligned copy grid (m0 postings), the others are whole.
          const int slack_l = (int)(e0 & 1);
          const int m0_l = min(cnt, kChunk - slack_l);
          const int nch = cnt > 0 ? 1 + (cnt - m0_l + kChunk - 1) / kChunk : 0;
          int incl = nch;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
          }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          for (int j0 = 0; j0 < total; j0 += kIssue) {      // rounds of kIssue chunks, one chunk per lane
            const int j = j0 + lane;
            int t = 0;                                      // term of chunk j = number of terms whose chunks end at or before j
            for (int u = 0; u < nt; ++u) t += (__shfl_sync(0xffffffffu, incl, u) <= j) ? 1 : 0;
            t = min(t, 31);
            const int incl_t = __shfl_sync(0xffffffffu, incl, t);
            const int nch_t = __shfl_sync(0xffffffffu, nch, t);
            const int cnt_t = __shfl_sync(0xffffffffu, cnt, t);
            const int64_t p_t = __shfl_sync(0xffffffffu, (long long)e0, t);
            const uint32_t w_t = __float_as_uint(__shfl_sync(0xffffffffu, wgt, t));
            const bool act = lane < kIssue && j < total;
            if (act) {
              const int c = j - (incl_t - nch_t);           // chunk index inside the term
              const int sl_t = (int)(p_t & 1);
              const int m0 = min(cnt_t, kChunk - sl_t);
              const int slack = c == 0 ? sl_t : 0;
              const int64_t p = c == 0 ? p_t : p_t + m0 + (int64_t)(c - 1) * kChunk;
              const int m = c == 0 ? m0 : min(kChunk, cnt_t - m0 - (c - 1) * kChunk);
              uint32_t fl = 0;
              if (first && j == 0) fl |= kFNewUnit;
              if (c == 0 && t != first_live) fl |= kFSync;
              if (j == total - 1) fl |= kFEndSpan;
              uint32_t st = s + (uint32_t)lane, php = ph;
              if (st >= (uint32_t)kStages) { st -= kStages; php ^= 1u; }
              mbar_wait(empty_u + st * 8u, php ^ 1u, a.status, 472);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(desc_u + st * 16u),
                           "r"((uint32_t)m | ((uint32_t)slack << 16) | (fl << 24)), "r"(w_t), "r"(doc0), "r"((uint32_t)unit)
                           : "memory");
              const uint32_t bytes = (uint32_t)((slack + m + 1) & ~1) * 8u;
              mbar_arrive_expect_tx(full_u + st * 8u, bytes);
              asm volatile(
                  "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                  :
                  : "r"(ring_u + st * (uint32_t)(kChunk * 8)), "l"(a.post + (p - slack)), "r"(bytes), "r"(full_u + st * 8u)
                  : "memory");
            }
            __syncwarp();
            s += (uint32_t)min(kIssue, total - j0);
            if (s >= (uint32_t)kStages) { s -= kStages; ph ^= 1u; }
          }
          first = false;
        }
        e0 = e1; e1 = e2; e2 = e3; e3 = e4;
      }
      put(((first ? kFNewUnit : 0u) | kFEndUnit) << 24, 0u, 0u, (uint32_t)unit, nullptr, 0u);
    }
    put(kFExit << 24, 0u, 0u, 0u, nullptr, 0u);
    return;
  }

  // ================= consumers =================
  const uint32_t acc_u = smem_u32(acc);
  volatile int* v_cnt = &s_int[2];
  volatile int* v_ovf = &s_int[3];
  float tau = 0.f;
  int unit = -1;
  int want = -1;      // tag filter of the current unit's query (< 0: none)
  uint32_t s = 0, ph = 0;
  auto compact = [&](int n) {
    const uint64_t T = block_compact_topk_t<kSpanThreads, kSpanCap, kSpanBar>(cand, n, a.k, hist, s_prefix, &s_int[0],
                                                                              &s_int[1], tid);
    // one ulp below the k-th best score: a later doc that ties with it but has a smaller id must still pass "> tau"
    tau = f32_from_orderable((uint32_t)(T >> 32) - 1u);
    if (tid == 0) { s_int[2] = s_int[1]; s_int[3] = 0; }
    bar_group<kSpanThreads, kSpanBar>();
  };
  for (;;) {
    // Every consumer warp passes here once per chunk, most of them without a posting of their own (a chunk holds
    // a few hundred postings on average): the common path is kept to the poll, the descriptor, the term barrier
    // and the release.
    if (!mbar_try_wait(full_u + s * 8u, ph)) mbar_wait(full_u + s * 8u, ph, a.status, 471);
    uint32_t dx, dy, dz, dw;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(dx), "=r"(dy), "=r"(dz), "=r"(dw) : "r"(desc_u + s * 16u));
    const uint32_t fl = dx >> 24;
    const int count = (int)(dx & 0xffffu);
    if (fl & kFSync) bar_group<kSpanThreads, kSpanBar>();     // the previous term's adds are complete
    if (tid < count) {
      const float w = __uint_as_float(dy);
      const uint32_t acc0 = acc_u - dz * 4u;                  // &acc[doc - doc0] == acc0 + doc * 4
      const uint32_t pa = ring_u + s * (uint32_t)(kChunk * 8) + (((dx >> 16) & 1u) + (uint32_t)tid) * 8u;
      uint32_t d[kPerThread];
      float im[kPerThread], o[kPerThread];
#pragma unroll
      for (int u = 0; u < kPerThread; ++u) {
        d[u] = 0; im[u] = 0.f;
        if (u == 0 || tid + u * kSpanThreads < count)
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(d[u]), "=f"(im[u]) : "r"(pa + (uint32_t)u * (kSpanThreads * 8u)));
      }
#pragma unroll
      for (int u = 0; u < kPerThread; ++u) {
        o[u] = 0.f;
        if (u == 0 || tid + u * kSpanThreads < count) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(o[u]) : "r"(acc0 + d[u] * 4u));
      }
#pragma unroll
      for (int u = 0; u < kPerThread; ++u)
        if (u == 0 || tid + u * kSpanThreads < count)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(acc0 + d[u] * 4u), "f"(__fadd_rn(o[u], __fmul_rn(w, im[u]))) : "memory");
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_u + s * 8u);
    if (++s == kStages) { s = 0; ph ^= 1u; }
    if (!(fl & (kFNewUnit | kFEndSpan | kFEndUnit | kFExit))) continue;
    if (fl & kFNewUnit) {
      unit = (int)dw;
      tau = 0.f;
      want = (a.tags && a.want) ? a.want[a.units[unit].q] : -1;
    }

    if (fl & kFEndSpan) {
      bar_group<kSpanThreads, kSpanBar>();                    // every add of the span has landed
      for (;;) {
        const uint32_t sa = acc_u + (uint32_t)tid * 16u;
        const uint32_t tau_u = tau > 0.f ? __float_as_uint(tau) : 0u;
        // groups of kScanBatch 16-byte loads in flight, then the tests (the loads of a group do not wait for the
        // stores of the previous slot)
#pragma unroll 1
        for (int j0 = 0; j0 < kScanIters; j0 += kScanBatch) {
          uint32_t vv[kScanBatch][4];
#pragma unroll
          for (int u = 0; u < kScanBatch; ++u)
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(vv[u][0]), "=r"(vv[u][1]), "=r"(vv[u][2]), "=r"(vv[u][3])
                         : "r"(sa + (uint32_t)(j0 + u) * (kSpanThreads * 16u)));
#pragma unroll
          for (int u = 0; u < kScanBatch; ++u) {
          const int j = j0 + u;
          const uint32_t addr = sa + (uint32_t)j * (kSpanThreads * 16u);
          const uint32_t(&v)[4] = vv[u];
          // Scores are >= 0, so their bit patterns order like unsigned integers: one integer max of the four slots
          // decides "all zero" (nothing to do), "none above tau" (zero the slots) or the rare append path, which
          // repeats the test exactly in fp32 (a negative score, possible only with a caller-made negative idf,
          // looks large here and is sorted out there).
          const uint32_t m = max(max(v[0], v[1]), max(v[2], v[3]));
          if (m > tau_u) {
            uint32_t z[4] = {0u, 0u, 0u, 0u};
            const uint32_t doc = dz + ((uint32_t)j * kSpanThreads + (uint32_t)tid) * 4u;
            // A doc outside the query's tag is dropped here: it never enters the list, so tau is learnt from
            // eligible docs only and the result is the exact top-k of the filtered corpus.  One 8-byte load brings
            // the tags of all four slots (doc is a multiple of 4).
            int tg[4] = {want, want, want, want};
            if (want >= 0) {
              if ((int64_t)doc + 4 <= a.n_docs) {
                const ushort4 t4 = __ldg((const ushort4*)(a.tags + doc));
                tg[0] = t4.x; tg[1] = t4.y; tg[2] = t4.z; tg[3] = t4.w;
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) tg[c] = (int64_t)doc + c < a.n_docs ? (int)a.tags[doc + c] : -1;
              }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float x = __uint_as_float(v[c]);
              if (x > tau && tg[c] == want) {
                const int pos = atomicAdd(&s_int[2], 1);
                if (pos < kSpanCap) cand[pos] = pack_key(x, doc + c);
                else { *v_ovf = 1; z[c] = v[c]; }              // stays in the accumulator for the next scan
              }
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3])
                         : "memory");
          } else if (m != 0u) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
          }
          }
        }
        bar_group<kSpanThreads, kSpanBar>();
        const int n = min(*v_cnt, kSpanCap);
        const bool ovf = *v_ovf != 0;
        // Under a tag filter every doc above tau costs a tag load whether it is eligible or not, so tau has to
        // follow the eligible docs closely: compact as soon as the list holds a few more than k entries.
        const int limit = want >= 0 ? min(kSpanCap / 2, max(2 * a.k, 256)) : kSpanCap / 2;
        if (!ovf && n <= limit) break;                        // uniform: the counters are stable here
        bar_group<kSpanThreads, kSpanBar>();                  // everyone has read them
        if (n > a.k) compact(n);
        else if (tid == 0) s_int[3] = 0;                      // (cannot overflow with n <= k; keep the flag sane)
        if (!ovf) break;
      }
    }
    if (fl & kFEndUnit) {
      bar_group<kSpanThreads, kSpanBar>();
      int n = min(*v_cnt, kSpanCap);
      bar_group<kSpanThreads, kSpanBar>();
      if (n > a.k) {
        (void)block_compact_topk_t<kSpanThreads, kSpanCap, kSpanBar>(cand, n, a.k, hist, s_prefix, &s_int[0], &s_int[1], tid);
        n = s_int[1];
        bar_group<kSpanThreads, kSpanBar>();
      }
      // bitonic sort (descending) of <= 256 keys padded with 0
      for (int i = n + tid; i < kMaxSelB; i += kSpanThreads) cand[i] = 0ull;
      bar_group<kSpanThreads, kSpanBar>();
      for (int size = 2; size <= kMaxSelB; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          if (tid < kMaxSelB / 2) {
            const int lo = ((tid / stride) * (stride << 1)) + (tid % stride);
            const int hi = lo + stride;
            const bool desc_block = ((lo & size) == 0);
            const uint64_t x = cand[lo], y = cand[hi];
            const bool swap = desc_block ? (y > x) : (x > y);
            if (swap) { cand[lo] = y; cand[hi] = x; }
          }
          bar_group<kSpanThreads, kSpanBar>();
        }
      }
      if (tid == 0) a.part_cnt[unit] = n;
      for (int i = tid; i < n; i += kSpanThreads) a.part_keys[(size_t)unit * a.k + i] = cand[i];
      bar_group<kSpanThreads, kSpanBar>();
      if (tid == 0) { s_int[2] = 0; s_int[3] = 0; }
      bar_group<kSpanThreads, kSpanBar>();
    }
    if (fl & kFExit) break;
  }
}

__global__ void bm25_df_kernel(const int64_t* skip, int n_blk, int V, int64_t* df) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  df[t] = skip[(size_t)(t + 1) * n_blk] - skip[(size_t)t * n_blk];
}

// cost[q] = total postings of the query's terms + term_cost per (term with postings, range): the kernel's time
// follows the number of (term, span) visits at least as much as the number of postings.
__global__ void bm25_cost_kernel(const int32_t* q_terms, const int32_t* q_off, const int64_t* df, int V,
                                 int B, int n_blk, long long term_cost, unsigned long long* keys,
                                 thr_dev_status* status) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  // more terms than the kernel has lanes for: reported by thr_sync, never silently truncated
  if (q_off[q + 1] - q_off[q] > kMaxTerms) dev_report(status, THR_EINVAL, 460, q);
  long long c = 0;
  for (int i = q_off[q]; i < q_off[q + 1]; ++i) {
    int t = q_terms[i];
    if (t >= 0 && t < V && df[t] > 0) c += df[t] + term_cost * n_blk;
  }
  if (c > 0xffffffffll) c = 0xffffffffll;
  keys[q] = ((unsigned long long)c << 32) | (unsigned)(0xffffffffu - (unsigned)q);
}

// Single block: cut queries into units of roughly equal cost.  A range costs its postings plus a fixed
// per-range overhead (kRangeCost postings' worth of pipeline work), so light queries are split as well.
constexpr unsigned long long kSpanRangeCost = 400;   // the scan of a range is worth about this many postings
constexpr long long kSpanTermCost = 75;         // per (term, range) on top of the term's postings
constexpr int kSpanUnitsPerCta = 2;
__global__ void __launch_bounds__(1024) bm25_plan_kernel(const unsigned long long* keys, int B, int n_blk,
                                                          int num_slots, unsigned long long kRangeCost,
                                                          Unit* units, int* unit_base,
                                                          int* total_units, int* work_counter) {
  __shared__ unsigned long long s_tot;
  __shared__ int s_carry;
  __shared__ int s_scan[1024];
  const int tid = threadIdx.x;
  if (tid == 0) { s_tot = 0; s_carry = 0; *work_counter = 0; }
  __syncthreads();
  unsigned long long part = 0;
  for (int q = tid; q < B; q += 1024) part += (keys[q] >> 32) + kRangeCost * (unsigned long long)n_blk;
  atomicAdd(&s_tot, part);
  __syncthreads();
  unsigned long long target = s_tot / (unsigned long long)num_slots + 1;
  if (target < 65536ull) target = 65536ull;
  for (int q0 = 0; q0 < B; q0 += 1024) {
    const int q = q0 + tid;
    int nu = 0;
    unsigned long long c = 0;
    if (q < B) {
      c = (keys[q] >> 32) + kRangeCost * (unsigned long long)n_blk;
      nu = (int)((c + target - 1) / target);
      if (nu < 1) nu = 1;
      if (nu > kMaxUnitsPerQuery) nu = kMaxUnitsPerQuery;
      if (nu > n_blk) nu = n_blk;
    }
    s_scan[tid] = nu;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {  // inclusive Hillis-Steele scan
      int v = tid >= off ? s_scan[tid - off] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int base = s_carry + s_scan[tid] - nu;
    if (q < B) {
      unit_base[q] = base;
      for (int u = 0; u < nu; ++u) {
        Unit x;
        x.q = q;
        x.r0 = (int)((long long)n_blk * u / nu);
        x.r1 = (int)((long long)n_blk * (u + 1) / nu);
        x.cost = (unsigned)min(c / (unsigned long long)nu, 0xffffffffull);
        units[base + u] = x;
      }
    }
    __syncthreads();
    if (tid == 1023) s_carry += s_scan[1023];
    __syncthreads();
  }
  if (tid == 0) { unit_base[B] = s_carry; *total_units = s_carry; }
}

// Rank sort of the units by cost, heaviest first (n <= B * kMaxUnitsPerQuery, a few thousand).
__global__ void bm25_order_kernel(const Unit* units, const int* total_units, int32_t* order) {
  const int n = *total_units;
  for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n; u += gridDim.x * blockDim.x) {
    const unsigned cu = units[u].cost;
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const unsigned cj = units[j].cost;
      rank += (cj > cu || (cj == cu && j < u)) ? 1 : 0;
    }
    order[rank] = u;
  }
}

// Per query: merge the sorted partial lists of its units -> final top-k.
__global__ void __launch_bounds__(256) bm25_merge_kernel(const uint64_t* part_keys, const int32_t* part_cnt,
                                                         const int* unit_base, int k, int64_t id_base,
                                                         int64_t* out_ids, float* out_scores, int32_t* out_count) {
  extern __shared__ uint64_t keys[];   // next power of two >= kMaxUnitsPerQuery * k
  const int q = blockIdx.x, tid = threadIdx.x;
  const int u0 = unit_base[q], u1 = unit_base[q + 1];
  const int slots = (u1 - u0) * k;
  int P = 32;
  while (P < slots) P <<= 1;
  for (int i = tid; i < P; i += 256) {
    uint64_t key = 0ull;
    if (i < slots) {
      const int u = u0 + i / k, j = i % k;
      if (j < part_cnt[u]) key = part_keys[(size_t)u * k + j];
    }
    keys[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= P; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (P >> 1); i += 256) {
        int lo = ((i / stride) * (stride << 1)) + (i % stride);
        int hi = lo + stride;
        bool desc_block = ((lo & size) == 0);
        uint64_t x = keys[lo], y = keys[hi];
        bool swap = desc_block ? (y > x) : (x > y);
        if (swap) { keys[lo] = y; keys[hi] = x; }
      }
      __syncthreads();
    }
  }
  int n = 0;
  for (int u = u0; u < u1; ++u) n += part_cnt[u];
  if (n > k) n = k;
  if (tid == 0) out_count[q] = n;
  for (int i = tid; i < k; i += 256) {
    size_t o = (size_t)q * k + i;
    if (i < n) {
      out_ids[o] = id_base + (int64_t)key_index(keys[i]);
      out_scores[o] = key_score(keys[i]);
    } else {
      out_ids[o] = -1;
      out_scores[o] = 0.f;
    }
  }
}

}  // namespace

struct thr_bm25_state {
  const int64_t* skip;
  const void* post;
  const float* idf;
  int64_t n_docs;
  int n_blk, blk_docs, blk_shift, V;
  int64_t id_base;
  int64_t* df;  // [V] device
  const uint16_t* tags;  // [n_docs] device, nullable
};

void thr_bm25_state_free(thr_handle* h) {
  if (h->bm25) {
    if (h->bm25->df) cudaFree(h->bm25->df);
    free(h->bm25);
    h->bm25 = nullptr;
  }
}

extern "C" {

int thr_bm25_index_set(thr_handle* h, const int64_t* skip, const void* postings, const float* idf,
                       int64_t n_docs, int32_t n_blk, int32_t blk_docs, int32_t V, int64_t id_base) {
  if (!h) return THR_EINVAL;
  cudaSetDevice(h->device);
  THR_REQUIRE(h, skip && postings && idf, "thr_bm25_index_set: NULL argument");
  THR_REQUIRE(h, n_docs >= 1 && V >= 1 && n_blk >= 1, "thr_bm25_index_set: empty index");
  int shift = 0;
  while ((1 << shift) < blk_docs) ++shift;
  if ((1 << shift) != blk_docs || blk_docs < 256 || blk_docs > kMaxBlkDocs)
    return thr_fail(h, THR_EUNSUPPORTED, "thr_bm25_index_set: blk_docs = %d must be a power of two in [256, %d]",
                    blk_docs, kMaxBlkDocs);
  THR_REQUIRE(h, (int64_t)n_blk * blk_docs >= n_docs && (int64_t)(n_blk - 1) * blk_docs < n_docs,
              "thr_bm25_index_set: n_blk does not match n_docs / blk_docs");
  THR_REQUIRE(h, n_docs < ((int64_t)1 << 32), "thr_bm25_index_set: more than 2^32 docs per shard");
  THR_REQUIRE(h, ((uintptr_t)postings & 15u) == 0, "thr_bm25_index_set: postings must be 16-byte aligned");
  THR_REQUIRE(h, ((uintptr_t)skip & 7u) == 0, "thr_bm25_index_set: skip must be 8-byte aligned");
  thr_bm25_state_free(h);
  thr_bm25_state* st = (thr_bm25_state*)calloc(1, sizeof(thr_bm25_state));
  if (!st) return thr_fail(h, THR_ENOMEM, "out of host memory");
  st->skip = skip; st->post = postings; st->idf = idf; st->n_docs = n_docs; st->n_blk = n_blk;
  st->blk_docs = blk_docs; st->blk_shift = shift; st->V = V; st->id_base = id_base;
  cudaError_t e = cudaMalloc((void**)&st->df, (size_t)V * sizeof(int64_t));
  if (e != cudaSuccess) { free(st); return thr_fail(h, THR_ENOMEM, "cudaMalloc(df): %s", cudaGetErrorString(e)); }
  bm25_df_kernel<<<(V + 255) / 256, 256>>>(skip, n_blk, V, st->df);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(st->df); free(st);
    return thr_fail(h, THR_ECUDA, "bm25_df_kernel: %s", cudaGetErrorString(e));
  }
  h->launches++;
  h->bm25 = st;
  return THR_OK;
}

int thr_bm25_tags_set(thr_handle* h, const uint16_t* tags) {
  if (!h) return THR_EINVAL;
  if (!h->bm25) return thr_fail(h, THR_ENOINDEX, "thr_bm25_tags_set: call thr_bm25_index_set first");
  THR_REQUIRE(h, ((uintptr_t)tags & 7u) == 0, "thr_bm25_tags_set: tags must be 8-byte aligned");
  h->bm25->tags = tags;
  return THR_OK;
}

int thr_bm25_topk(thr_handle* h, const int32_t* q_terms, const int32_t* q_off, int B, int k,
                  int64_t* out_ids, float* out_scores, int32_t* out_count, void* stream) {
  return thr_bm25_topk_tagged(h, q_terms, q_off, B, k, nullptr, out_ids, out_scores, out_count, stream);
}

int thr_bm25_topk_tagged(thr_handle* h, const int32_t* q_terms, const int32_t* q_off, int B, int k,
                         const int32_t* want, int64_t* out_ids, float* out_scores, int32_t* out_count,
                         void* stream) {
  if (!h) return THR_EINVAL;
  cudaSetDevice(h->device);
  thr_bm25_state* st = h->bm25;
  if (!st) return thr_fail(h, THR_ENOINDEX, "thr_bm25_topk: call thr_bm25_index_set first");
  THR_REQUIRE(h, want == nullptr || st->tags != nullptr, "thr_bm25_topk_tagged: call thr_bm25_tags_set first");
  THR_REQUIRE(h, B >= 0 && k >= 1 && k <= kMaxSelB, "thr_bm25_topk: need 1 <= k <= %d", kMaxSelB);
  if (B == 0) return THR_OK;
  THR_REQUIRE(h, q_terms && q_off && out_ids && out_scores && out_count, "thr_bm25_topk: NULL argument");
  cudaStream_t s = (cudaStream_t)stream;
  // scratch: cost keys | units | order | unit_base | counters | partial lists
  const size_t max_units = (size_t)B * kMaxUnitsPerQuery;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_keys = 0;
  const size_t o_units = o_keys + up((size_t)B * 8);
  const size_t o_order = o_units + up(max_units * sizeof(Unit));
  const size_t o_base = o_order + up(max_units * 4);
  const size_t o_cnt = o_base + up((size_t)(B + 1) * 4);
  const size_t o_pcnt = o_cnt + 256;
  const size_t o_pkeys = o_pcnt + up(max_units * 4);
  const size_t need = o_pkeys + up(max_units * (size_t)k * 8);
  uint8_t* ws = (uint8_t*)thr_scratch(h, need);
  if (!ws) return THR_ENOMEM;
  unsigned long long* keys = (unsigned long long*)(ws + o_keys);
  Unit* units = (Unit*)(ws + o_units);
  int32_t* order = (int32_t*)(ws + o_order);
  int* unit_base = (int*)(ws + o_base);
  int* counter = (int*)(ws + o_cnt);
  int* total_units = counter + 1;
  int32_t* part_cnt = (int32_t*)(ws + o_pcnt);
  uint64_t* part_keys = (uint64_t*)(ws + o_pkeys);

  int tok = thr_prof_begin(h, THR_PROF_BM25_PREP, s);

  // Work is cut into about `grid * units_per_cta` units of equal cost (heaviest first, fetched dynamically).
  static int units_per_cta = 0;
  static long long range_cost = -1, term_cost = 0;
  if (!units_per_cta) {
    const char* t = getenv("THR_BM25_TERM_COST");
    term_cost = t ? atoll(t) : (long long)kSpanTermCost;
    const char* e = getenv("THR_BM25_UNITS_PER_SM");   // measured flat between 1 and 4 at 10M docs
    units_per_cta = e ? atoi(e) : kSpanUnitsPerCta;
    if (units_per_cta < 1) units_per_cta = 1;
    e = getenv("THR_BM25_RANGE_COST");
    range_cost = e ? atoll(e) : (long long)kSpanRangeCost;
  }
  const int grid = kSpanCtas * h->num_sms;
  bm25_cost_kernel<<<(B + 255) / 256, 256, 0, s>>>(q_terms, q_off, st->df, st->V, B, st->n_blk, term_cost, keys,
                                                   h->d_status);
  THR_CHECK_LAUNCH(h, "bm25_cost_kernel");
  bm25_plan_kernel<<<1, 1024, 0, s>>>(keys, B, st->n_blk, grid * units_per_cta, (unsigned long long)range_cost, units,
                                      unit_base, total_units, counter);
  THR_CHECK_LAUNCH(h, "bm25_plan_kernel");
  bm25_order_kernel<<<32, 256, 0, s>>>(units, total_units, order);
  thr_prof_end(h, tok, s);
  THR_CHECK_LAUNCH(h, "bm25_order_kernel");

  Bm25Args a;
  a.skip = st->skip; a.post = (const Posting*)st->post; a.idf = st->idf; a.n_docs = st->n_docs;
  a.n_blk = st->n_blk; a.blk_docs = st->blk_docs; a.blk_shift = st->blk_shift; a.V = st->V;
  a.q_terms = q_terms; a.q_off = q_off; a.order = order; a.units = units; a.total_units = total_units;
  a.work_counter = counter; a.B = B; a.k = k; a.part_keys = part_keys; a.part_cnt = part_cnt;
  a.tags = want ? st->tags : nullptr; a.want = want;
  a.status = h->d_status;
  THR_CUDA(h, cudaFuncSetAttribute(bm25_span_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSpanSmem));
  THR_CUDA(h, cudaFuncSetAttribute(bm25_span_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  tok = thr_prof_begin(h, THR_PROF_BM25, s);
  bm25_span_kernel<<<grid, kSpanThreads + 32, kSpanSmem, s>>>(a);
  thr_prof_end(h, tok, s);
  THR_CHECK_LAUNCH(h, "bm25_span_kernel");
  tok = thr_prof_begin(h, THR_PROF_BM25_PREP, s);
  size_t merge_slots = 32;
  while (merge_slots < (size_t)kMaxUnitsPerQuery * k) merge_slots <<= 1;
  THR_CUDA(h, cudaFuncSetAttribute(bm25_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(merge_slots * 8)));
  bm25_merge_kernel<<<B, 256, merge_slots * 8, s>>>(part_keys, part_cnt, unit_base, k, st->id_base, out_ids, out_scores,
                                                    out_count);
  thr_prof_end(h, tok, s);
  THR_CHECK_LAUNCH(h, "bm25_merge_kernel");
  return THR_OK;
}

}  // extern "C"
