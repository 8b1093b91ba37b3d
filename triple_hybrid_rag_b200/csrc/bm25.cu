// bm25.cu — K2: BM25 top-k over a blocked CSR inverted index (lexical channel).
//
// Stands where RAG2Retriever._lexical_search calls the rag2_lexical_search RPC
//   (src/voice_agent/rag2/retrieval.py:273-292, database/migrations/20260114_rag2_schema.sql:341-374):
//   same interface (keyword list in, top-`limit` by descending score out); the scoring formula is
//   BM25 as BASELINE.json's north_star asks.  Definition (oracle/bm25.py restates it on the CPU):
//     score(d) = fp32 sum, over the query's terms in order, of  idf[t] * impact(t, d)
//   with fp32 round-to-nearest multiply and add and no FMA contraction, so the result does not
//   depend on scheduling; docs with score > 0 are ranked by (score desc, id asc).
//
// One CTA works on one query at a time (persistent grid, queries handed out heaviest first).
// The doc space is walked range by range (blk_docs docs, fp32 accumulators in shared memory).
// Warp 16 is the producer: for every range it looks up the query terms' posting segments
// (blk_ptr), and moves them global -> shared with cp.async.bulk into a 3-stage ring, completion on
// an mbarrier.  Warps 0-15 consume: per term a coalesced pass shared -> accumulator (doc ids are
// unique inside a posting list, terms are separated by a named barrier => no atomics).  The posting
// that touches an accumulator first marks itself as the doc's owner (bit 31 of the staged doc id);
// a second pass over the staged postings lets every owner read its doc's final score, append it to
// the candidate list if it beats the running threshold (warp-aggregated) and re-zero the slot — so
// the work per range is proportional to its postings, not to blk_docs.  Ranges too large for one
// ring stage fall back to a scan of all accumulators.  The candidate list is compacted to the best k
// by a block radix select when it fills; that also raises the threshold.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kMaxTerms = 32;
constexpr int kConsumerWarps = 16;
constexpr int kConsumers = kConsumerWarps * 32;  // 512
constexpr int kThreads = kConsumers + 32;        // + producer warp
constexpr int kMaxBlkDocs = 16384;
constexpr int kStageCap = 4096;                  // postings per ring stage (32 KB)
constexpr int kStages = 3;
constexpr int kCandCap = 4608;                   // >= kStageCap + kMaxSelB (one sparse pass can append a whole stage)
constexpr int kMaxSelB = 256;
constexpr int kScanChunk = 2048;                 // accumulator slots scanned between capacity checks

struct Posting { uint32_t doc; float imp; };

// A work unit: one query restricted to the doc ranges [r0, r1).  Heavy queries are cut into several
// units so that no CTA is left streaming one long query while the others idle.
struct Unit { int q; int r0; int r1; unsigned cost; };
constexpr int kMaxUnitsPerQuery = 16;

struct StageMeta {
  int nseg;
  int last_of_range;     // scan after this step
  int range;             // range index
  int range_add_bound;   // upper bound of docs this range can append (valid on last step)
  int end_of_query;      // no data: consumers finish the query
  int single;            // this step holds ALL postings of the range -> sparse second pass
  int used;              // postings in this step (sum of seg_count)
  int seg_term[kMaxTerms + 2];    // query term slot
  int seg_start[kMaxTerms + 2];   // first valid posting inside the stage buffer
  int seg_count[kMaxTerms + 2];
  int pad_;
};
static_assert(sizeof(StageMeta) % 8 == 0, "mbarriers follow the metadata and need 8-byte alignment");

struct Bm25Args {
  const int64_t* blk_ptr;
  const Posting* post;
  const float* idf;
  int64_t n_docs;
  int n_blk, blk_docs, V;
  int64_t id_base;
  const int32_t* q_terms;
  const int32_t* q_off;
  const int32_t* order;   // work units, heaviest first
  const Unit* units;
  const int* total_units;
  int* work_counter;
  int B, k;
  uint64_t* part_keys;    // [max_units][k] sorted descending
  int32_t* part_cnt;      // [max_units]
  thr_dev_status* status;
};

__device__ __forceinline__ void bar_consumers() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// Block-cooperative (consumer threads only): keep the ksel largest of keys[0..n) in place, n > ksel.
// Returns the ksel-th largest key.  hist/scal are shared scratch.
__device__ uint64_t block_compact_topk(uint64_t* keys, int n, int ksel, uint32_t* hist,
                                       unsigned long long* s_prefix, int* s_want, int* s_cnt, int tid) {
  if (tid == 0) { *s_prefix = 0ull; *s_want = ksel; }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    bar_consumers();
    const uint64_t prefix = *s_prefix;
    for (int i = tid; i < n; i += kConsumers) {
      const uint64_t key = keys[i];
      const bool match = pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8));
      if (match) atomicAdd(&hist[(uint32_t)(key >> shift) & 255u], 1u);
    }
    bar_consumers();
    if (tid == 0) {
      int want = *s_want, cum = 0, d = 255;
      for (; d > 0; --d) {
        if (cum + (int)hist[d] >= want) break;
        cum += hist[d];
      }
      *s_want = want - cum;
      *s_prefix = prefix | ((unsigned long long)d << shift);
    }
    bar_consumers();
  }
  const uint64_t T = *s_prefix;
  // survivors: read everything first, then rewrite the front
  constexpr int kPer = kCandCap / kConsumers;
  uint64_t mine[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    int i = tid + j * kConsumers;
    mine[j] = i < n ? keys[i] : 0ull;
  }
  if (tid == 0) *s_cnt = 0;
  bar_consumers();
#pragma unroll
  for (int j = 0; j < kPer; ++j)
    if (mine[j] >= T && mine[j] != 0ull) keys[atomicAdd(s_cnt, 1)] = mine[j];
  bar_consumers();
  return T;
}

__global__ void __launch_bounds__(kThreads, 1) bm25_kernel(const Bm25Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // carve-up
  Posting* stage_buf = (Posting*)gen;                                         // kStages * kStageCap
  float* acc = (float*)(gen + (size_t)kStages * kStageCap * sizeof(Posting)); // kMaxBlkDocs
  uint64_t* cand = (uint64_t*)(acc + kMaxBlkDocs);                            // kCandCap
  StageMeta* meta = (StageMeta*)(cand + kCandCap);                            // kStages
  uint64_t* bars = (uint64_t*)(meta + kStages);                               // full[kStages], empty[kStages]
  uint32_t* hist = (uint32_t*)(bars + 2 * kStages);                           // 256
  int* q_term = (int*)(hist + 256);                                           // kMaxTerms
  float* q_idf = (float*)(q_term + kMaxTerms);                                // kMaxTerms
  unsigned long long* s_prefix = (unsigned long long*)(q_idf + kMaxTerms);
  int* s_int = (int*)(s_prefix + 1);  // [0]=want [1]=cnt(compact) [2]=cand count [3]=query [4]=nterms

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[kStages + s]); };

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kMaxBlkDocs; i += kThreads) acc[i] = 0.f;
  __syncthreads();

  uint32_t it = 0;  // ring step counter, advances identically in producer and consumers
  const int R = a.blk_docs;

  for (;;) {
    // ---- fetch the next query (whole CTA) ----
    __syncthreads();
    if (tid == 0) {
      int w = atomicAdd(a.work_counter, 1);
      s_int[3] = w < *a.total_units ? a.order[w] : -1;
    }
    __syncthreads();
    const int unit = s_int[3];
    if (unit < 0) break;
    const int q = a.units[unit].q;
    const int r_begin = a.units[unit].r0, r_end = a.units[unit].r1;
    if (tid < kMaxTerms) {
      const int lo = a.q_off[q], hi = a.q_off[q + 1];
      int nt = hi - lo;
      if (nt > kMaxTerms) nt = kMaxTerms;  // host rejects longer queries; keep the kernel safe
      int t = -1;
      float w = 0.f;
      if (tid < nt) {
        t = a.q_terms[lo + tid];
        if (t < 0 || t >= a.V) t = -1; else w = a.idf[t];
      }
      q_term[tid] = t;
      q_idf[tid] = w;
      if (tid == 0) { s_int[4] = nt; s_int[2] = 0; }
    }
    __syncthreads();
    const int nterms = s_int[4];

    if (warp == kConsumerWarps) {
      // ======================= producer warp =======================
      const int my_term = lane < nterms ? q_term[lane] : -1;
      // pointer pipeline: the (lo, hi) pair of range r + 4 is requested while range r is packed
      auto load_ptr = [&](int r, int64_t& lo, int64_t& hi) {
        lo = 0; hi = 0;
        if (my_term >= 0 && r < r_end) {
          const int64_t* p = a.blk_ptr + (size_t)r * (a.V + 1) + my_term;
          lo = __ldg(p);
          hi = __ldg(p + 1);
        }
      };
      int64_t l0, h0, l1, h1, l2, h2, l3, h3;
      load_ptr(r_begin, l0, h0); load_ptr(r_begin + 1, l1, h1); load_ptr(r_begin + 2, l2, h2);
      load_ptr(r_begin + 3, l3, h3);
      for (int r = r_begin; r < r_end; ++r) {
        int64_t lo = l0, hi = h0;
        l0 = l1; h0 = h1; l1 = l2; h1 = h2; l2 = l3; h2 = h3;
        load_ptr(r + 4, l3, h3);
        int64_t remaining = hi - lo;
        // upper bound on distinct docs this range can contribute
        long long tot = remaining;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
        const int add_bound = (int)min((long long)R, tot);
        if (tot == 0) continue;  // nothing in this range for this query
        bool first_step = true;
        {
          // Fast path (the common case): the whole range fits in one ring stage.  Lanes pack their
          // segments with one warp scan and every lane issues its own bulk copy — no serial loop.
          const int cnt_l = (int)remaining;
          const int slack_l = (int)(lo & 1);
          const int cp_l = cnt_l > 0 ? ((slack_l + cnt_l + 1) & ~1) : 0;  // postings copied (16 B granules)
          int incl = cp_l;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
          }
          const int total_cp = __shfl_sync(0xffffffffu, incl, 31);
          if (total_cp <= kStageCap) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u, a.status, 402);
            const unsigned have = __ballot_sync(0xffffffffu, cnt_l > 0);
            const int used_before = incl - cp_l;
            if (cnt_l > 0) {
              const int g = __popc(have & ((1u << lane) - 1));
              meta[s].seg_term[g] = lane;
              meta[s].seg_start[g] = used_before + slack_l;
              meta[s].seg_count[g] = cnt_l;
            }
            __syncwarp();
            if (lane == 0) {
              meta[s].nseg = __popc(have);
              meta[s].last_of_range = 1;
              meta[s].range = r;
              meta[s].range_add_bound = add_bound;
              meta[s].end_of_query = 0;
              meta[s].single = 1;
              meta[s].used = (int)tot;
              mbar_arrive_expect_tx(full_bar(s), (uint32_t)total_cp * 8u);
            }
            __syncwarp();
            if (cnt_l > 0)
              bulk_g2s(smem_u32(stage_buf + (size_t)s * kStageCap + used_before), a.post + (lo - slack_l),
                       (uint32_t)cp_l * 8u, full_bar(s));
            ++it;
            continue;
          }
        }
        // emit steps until every lane's segment is drained (terms in order)
        while (true) {
          unsigned live = __ballot_sync(0xffffffffu, remaining > 0);
          if (!live) break;
          const int cur = __ffs(live) - 1;
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u, a.status, 400);
          // greedy packing in term order; each segment is copied from its 16-byte-aligned start
          int used = 0;  // postings used in the stage (including alignment slack)
          int nseg = 0, npost = 0;
          uint32_t bytes_total = 0;
          for (int t = cur; t < nterms; ++t) {
            int64_t rem_t = __shfl_sync(0xffffffffu, remaining, t);
            int64_t lo_t = __shfl_sync(0xffffffffu, lo, t);
            if (rem_t <= 0) continue;
            const int slack = (int)(lo_t & 1);           // start is 8-byte aligned; copy from the even posting
            int room = kStageCap - used - slack - 1;     // -1: the copy length is rounded up to 2 postings
            if (room < 64 && nseg > 0) break;            // keep pieces reasonably long
            if (room <= 0) break;
            int take = (int)min((int64_t)room, rem_t);
            const int64_t src_first = lo_t - slack;
            const int cp_postings = (slack + take + 1) & ~1;
            if (lane == 0) {
              meta[s].seg_term[nseg] = t;
              meta[s].seg_start[nseg] = used + slack;
              meta[s].seg_count[nseg] = take;
            }
            if (lane == t) {
              bulk_g2s(smem_u32(stage_buf + (size_t)s * kStageCap + used), a.post + src_first,
                       (uint32_t)cp_postings * 8u, full_bar(s));
              lo += take;
              remaining -= take;
            }
            bytes_total += (uint32_t)cp_postings * 8u;
            used += cp_postings;
            npost += take;
            ++nseg;
            if (take < rem_t) break;  // stage full in the middle of this term
          }
          unsigned still = __ballot_sync(0xffffffffu, remaining > 0);
          if (lane == 0) {
            meta[s].nseg = nseg;
            meta[s].last_of_range = still ? 0 : 1;
            meta[s].range = r;
            meta[s].range_add_bound = add_bound;
            meta[s].end_of_query = 0;
            meta[s].single = (first_step && !still) ? 1 : 0;
            meta[s].used = npost;
            // metadata is written with generic stores; the arrive has release semantics
            mbar_arrive_expect_tx(full_bar(s), bytes_total);
          }
          first_step = false;
          ++it;
          __syncwarp();
        }
      }
      // end-of-query marker
      {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u, a.status, 401);
        if (lane == 0) {
          meta[s].nseg = 0;
          meta[s].last_of_range = 0;
          meta[s].end_of_query = 1;
          meta[s].single = 0;
          meta[s].used = 0;
          mbar_arrive(full_bar(s));
        }
        ++it;
        __syncwarp();
      }
    } else {
      // ======================= consumers =======================
      float tau = 0.f;  // only score > 0 is eligible; raised by compactions
      for (;;) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1u;
        mbar_wait(full_bar(s), ph, a.status, 410);
        ++it;
        const StageMeta& m = meta[s];
        const int nseg = m.nseg;
        const bool eoq = m.end_of_query != 0;
        const bool last = m.last_of_range != 0;
        const bool single = m.single != 0;
        const int used = m.used;
        const int r = m.range;
        const int add_bound = m.range_add_bound;
        Posting* sb = stage_buf + (size_t)s * kStageCap;
        const uint32_t doc0 = (uint32_t)r * (uint32_t)R;
        // ---- pass 1: accumulate, term by term ----
        int prev_term = -1;
        for (int g = 0; g < nseg; ++g) {
          const int t = m.seg_term[g];
          const int st = m.seg_start[g];
          const int cn = m.seg_count[g];
          if (prev_term >= 0 && t != prev_term) bar_consumers();  // term boundary: same doc may recur
          prev_term = t;
          const float w = q_idf[t];
          for (int i = tid; i < cn; i += kConsumers) {
            const Posting p = sb[st + i];
            const uint32_t slot = p.doc - doc0;
            const float old = acc[slot];
            acc[slot] = __fadd_rn(old, __fmul_rn(w, p.imp));
            if (single && old == 0.f) sb[st + i].doc = p.doc | 0x80000000u;  // first touch owns the doc
          }
        }
        if (!single) {
          // this stage's shared buffer is free once every consumer warp has read it
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(s));
        }
        if (eoq) break;
        if (nseg > 0) bar_consumers();  // all accumulates of this step done

        if (single) {
          // ---- pass 2 (sparse): owners read the final score, append if > tau, re-zero ----
          int cnt_now = s_int[2];
          if (cnt_now + used > kCandCap) {  // block-uniform; implies cnt_now > kMaxSelB >= k
            const uint64_t T = block_compact_topk(cand, cnt_now, a.k, hist, s_prefix, &s_int[0], &s_int[1], tid);
            tau = key_score(T);
            if (tid == 0) s_int[2] = s_int[1];
            bar_consumers();
          }
          for (int g = 0; g < nseg; ++g) {
            const int st = m.seg_start[g];
            const int cn = m.seg_count[g];
            for (int i0 = warp * 32; i0 < cn; i0 += kConsumers) {
              const int i = i0 + lane;
              bool emit = false;
              uint64_t key = 0;
              if (i < cn) {
                const uint32_t d = sb[st + i].doc;
                if (d & 0x80000000u) {
                  const uint32_t doc = d & 0x7fffffffu;
                  const uint32_t slot = doc - doc0;
                  const float v = acc[slot];
                  acc[slot] = 0.f;
                  emit = v > tau;
                  key = pack_key(v, doc);
                }
              }
              const unsigned bal = __ballot_sync(0xffffffffu, emit);
              if (bal) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&s_int[2], __popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (emit) cand[base + __popc(bal & ((1u << lane) - 1))] = key;
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(s));
          bar_consumers();  // slots re-zeroed and appends visible before the next range
          continue;
        }
        if (!last) continue;

        // ---- fallback for ranges spread over several steps: scan all accumulators ----
        const int ndocs_r = (int)min((int64_t)R, a.n_docs - (int64_t)r * R);
        int cnt_now = s_int[2];
        const bool tight = cnt_now + add_bound > kCandCap;  // block-uniform
        for (int c0 = 0; c0 < ndocs_r; c0 += kScanChunk) {
          if (tight) {
            bar_consumers();
            cnt_now = s_int[2];
            if (cnt_now + kScanChunk > kCandCap) {
              const uint64_t T = block_compact_topk(cand, cnt_now, a.k, hist, s_prefix, &s_int[0], &s_int[1], tid);
              tau = key_score(T);
              if (tid == 0) s_int[2] = s_int[1];
              bar_consumers();
            }
          }
          const int c1 = min(c0 + kScanChunk, ndocs_r);
          for (int i = c0 + tid * 4; i < c1; i += kConsumers * 4) {
            float4 v = *reinterpret_cast<float4*>(&acc[i]);
            const float vv[4] = {v.x, v.y, v.z, v.w};
            bool any = false;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (vv[e] != 0.f) any = true;
              if (vv[e] > tau && i + e < c1) {
                int pos = atomicAdd(&s_int[2], 1);
                if (pos < kCandCap) cand[pos] = pack_key(vv[e], doc0 + (uint32_t)(i + e));
                else dev_report(a.status, THR_EOVERFLOW, 420, pos);
              }
            }
            if (any) *reinterpret_cast<float4*>(&acc[i]) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        bar_consumers();  // scan complete before the next range accumulates
      }

      // ---- end of query: final top-k, sorted ----
      bar_consumers();
      int n = s_int[2];
      if (n > a.k) {
        (void)block_compact_topk(cand, n, a.k, hist, s_prefix, &s_int[0], &s_int[1], tid);
        n = s_int[1];
        bar_consumers();
      }
      // bitonic sort (descending) of <= 256 keys padded with 0
      for (int i = n + tid; i < kMaxSelB; i += kConsumers) cand[i] = 0ull;
      bar_consumers();
      for (int size = 2; size <= kMaxSelB; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          if (tid < kMaxSelB / 2) {
            int lo = ((tid / stride) * (stride << 1)) + (tid % stride);
            int hi = lo + stride;
            bool desc_block = ((lo & size) == 0);
            uint64_t x = cand[lo], y = cand[hi];
            bool swap = desc_block ? (y > x) : (x > y);
            if (swap) { cand[lo] = y; cand[hi] = x; }
          }
          bar_consumers();
        }
      }
      if (tid == 0) a.part_cnt[unit] = n;
      for (int i = tid; i < n; i += kConsumers) a.part_keys[(size_t)unit * a.k + i] = cand[i];
    }
  }
}

// cost[q] = total postings of the query's terms.
__global__ void bm25_df_kernel(const int64_t* blk_ptr, int n_blk, int V, int64_t* df) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  int64_t s = 0;
  for (int r = 0; r < n_blk; ++r) {
    const int64_t* p = blk_ptr + (size_t)r * (V + 1) + t;
    s += p[1] - p[0];
  }
  df[t] = s;
}

__global__ void bm25_cost_kernel(const int32_t* q_terms, const int32_t* q_off, const int64_t* df, int V,
                                 int B, unsigned long long* keys) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B) return;
  long long c = 0;
  for (int i = q_off[q]; i < q_off[q + 1]; ++i) {
    int t = q_terms[i];
    if (t >= 0 && t < V) c += df[t];
  }
  if (c > 0xffffffffll) c = 0xffffffffll;
  keys[q] = ((unsigned long long)c << 32) | (unsigned)(0xffffffffu - (unsigned)q);
}

// Single block: cut queries into units of roughly equal posting count.
__global__ void __launch_bounds__(1024) bm25_plan_kernel(const unsigned long long* keys, int B, int n_blk,
                                                          int num_sms, Unit* units, int* unit_base,
                                                          int* total_units, int* work_counter) {
  __shared__ unsigned long long s_tot;
  __shared__ int s_carry;
  __shared__ int s_scan[1024];
  const int tid = threadIdx.x;
  if (tid == 0) { s_tot = 0; s_carry = 0; *work_counter = 0; }
  __syncthreads();
  unsigned long long part = 0;
  for (int q = tid; q < B; q += 1024) part += keys[q] >> 32;
  atomicAdd(&s_tot, part);
  __syncthreads();
  unsigned long long target = s_tot / (unsigned long long)(num_sms * 2) + 1;
  if (target < 32768ull) target = 32768ull;
  for (int q0 = 0; q0 < B; q0 += 1024) {
    const int q = q0 + tid;
    int nu = 0;
    unsigned long long c = 0;
    if (q < B) {
      c = keys[q] >> 32;
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
  __shared__ uint64_t keys[kMaxUnitsPerQuery * kMaxSelB];
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

constexpr size_t kBm25Smem = (size_t)kStages * kStageCap * 8 + (size_t)kMaxBlkDocs * 4 +
                             (size_t)kCandCap * 8 + kStages * sizeof(StageMeta) + 2 * kStages * 8 +
                             256 * 4 + kMaxTerms * 8 + 8 + 8 * 4 + 256;

}  // namespace

struct thr_bm25_state {
  const int64_t* blk_ptr;
  const void* post;
  const float* idf;
  int64_t n_docs;
  int n_blk, blk_docs, V;
  int64_t id_base;
  int64_t* df;  // [V] device
};

void thr_bm25_state_free(thr_handle* h) {
  if (h->bm25) {
    if (h->bm25->df) cudaFree(h->bm25->df);
    free(h->bm25);
    h->bm25 = nullptr;
  }
}

extern "C" {

int thr_bm25_index_set(thr_handle* h, const int64_t* blk_ptr, const void* postings, const float* idf,
                       int64_t n_docs, int32_t n_blk, int32_t blk_docs, int32_t V, int64_t id_base) {
  if (!h) return THR_EINVAL;
  cudaSetDevice(h->device);
  THR_REQUIRE(h, blk_ptr && postings && idf, "thr_bm25_index_set: NULL argument");
  THR_REQUIRE(h, n_docs >= 1 && V >= 1 && n_blk >= 1, "thr_bm25_index_set: empty index");
  if (blk_docs % 1024 != 0 || blk_docs < 1024 || blk_docs > kMaxBlkDocs)
    return thr_fail(h, THR_EUNSUPPORTED, "thr_bm25_index_set: blk_docs = %d must be a multiple of 1024 in [1024, %d]",
                    blk_docs, kMaxBlkDocs);
  THR_REQUIRE(h, (int64_t)n_blk * blk_docs >= n_docs && (int64_t)(n_blk - 1) * blk_docs < n_docs,
              "thr_bm25_index_set: n_blk does not match n_docs / blk_docs");
  THR_REQUIRE(h, n_docs < ((int64_t)1 << 31), "thr_bm25_index_set: more than 2^31 docs per shard (bit 31 of a staged doc id is the owner flag)");
  THR_REQUIRE(h, ((uintptr_t)postings & 15u) == 0, "thr_bm25_index_set: postings must be 16-byte aligned");
  thr_bm25_state_free(h);
  thr_bm25_state* st = (thr_bm25_state*)calloc(1, sizeof(thr_bm25_state));
  if (!st) return thr_fail(h, THR_ENOMEM, "out of host memory");
  st->blk_ptr = blk_ptr; st->post = postings; st->idf = idf; st->n_docs = n_docs; st->n_blk = n_blk;
  st->blk_docs = blk_docs; st->V = V; st->id_base = id_base;
  cudaError_t e = cudaMalloc((void**)&st->df, (size_t)V * sizeof(int64_t));
  if (e != cudaSuccess) { free(st); return thr_fail(h, THR_ENOMEM, "cudaMalloc(df): %s", cudaGetErrorString(e)); }
  bm25_df_kernel<<<(V + 255) / 256, 256>>>(blk_ptr, n_blk, V, st->df);
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

int thr_bm25_topk(thr_handle* h, const int32_t* q_terms, const int32_t* q_off, int B, int k,
                  int64_t* out_ids, float* out_scores, int32_t* out_count, void* stream) {
  if (!h) return THR_EINVAL;
  cudaSetDevice(h->device);
  thr_bm25_state* st = h->bm25;
  if (!st) return thr_fail(h, THR_ENOINDEX, "thr_bm25_topk: call thr_bm25_index_set first");
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
  bm25_cost_kernel<<<(B + 255) / 256, 256, 0, s>>>(q_terms, q_off, st->df, st->V, B, keys);
  THR_CHECK_LAUNCH(h, "bm25_cost_kernel");
  bm25_plan_kernel<<<1, 1024, 0, s>>>(keys, B, st->n_blk, h->num_sms, units, unit_base, total_units, counter);
  THR_CHECK_LAUNCH(h, "bm25_plan_kernel");
  bm25_order_kernel<<<32, 256, 0, s>>>(units, total_units, order);
  thr_prof_end(h, tok, s);
  THR_CHECK_LAUNCH(h, "bm25_order_kernel");

  Bm25Args a;
  a.blk_ptr = st->blk_ptr; a.post = (const Posting*)st->post; a.idf = st->idf; a.n_docs = st->n_docs;
  a.n_blk = st->n_blk; a.blk_docs = st->blk_docs; a.V = st->V; a.id_base = st->id_base;
  a.q_terms = q_terms; a.q_off = q_off; a.order = order; a.units = units; a.total_units = total_units;
  a.work_counter = counter; a.B = B; a.k = k; a.part_keys = part_keys; a.part_cnt = part_cnt;
  a.status = h->d_status;
  THR_CUDA(h, cudaFuncSetAttribute(bm25_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBm25Smem));
  tok = thr_prof_begin(h, THR_PROF_BM25, s);
  bm25_kernel<<<h->num_sms, kThreads, kBm25Smem, s>>>(a);
  thr_prof_end(h, tok, s);
  THR_CHECK_LAUNCH(h, "bm25_kernel");
  tok = thr_prof_begin(h, THR_PROF_BM25_PREP, s);
  bm25_merge_kernel<<<B, 256, 0, s>>>(part_keys, part_cnt, unit_base, k, st->id_base, out_ids, out_scores, out_count);
  thr_prof_end(h, tok, s);
  THR_CHECK_LAUNCH(h, "bm25_merge_kernel");
  return THR_OK;
}

}  // extern "C"
