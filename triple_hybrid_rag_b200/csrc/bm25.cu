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
// One CTA works on one unit (a query restricted to a span of doc ranges; heavy queries are cut into
// several units) at a time, walking the span range by range with fp32 accumulators for one range in
// shared memory.  Two warp roles, asynchronous to each other through mbarriers:
//   producer (1 warp)  per range: reads the query terms' skip entries (prefetched 12 ranges ahead with
//                      cp.async), packs the terms' posting segments into one step, allocates room in a
//                      96 KB ring and moves the segments global -> shared with cp.async.bulk.
//   consumers          per step, term by term: one coalesced pass shared -> accumulator (doc ids are
//                      unique inside a posting list and terms are separated by a named barrier => no
//                      atomics).  The posting that touches an accumulator first marks itself as the
//                      doc's owner (bit 31 of the staged doc id); one more pass over the staged postings
//                      lets every owner read its doc's final score, append it to the candidate list if
//                      it beats the running threshold and re-zero the slot, so the work per range is
//                      proportional to its postings, not to blk_docs.  A range too large for one step
//                      falls back to a scan of the accumulators.
// The barrier after a range's last term doubles as the capacity vote (bar.red.or): when the candidate
// list could overflow it is compacted to the best k by a radix select, which also raises the threshold.
// The per-step code is kept to ~150 instructions per thread: the kernel is issue-bound, not DRAM-bound.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int kMaxTerms = 32;
#ifndef THR_BM25_CW
#define THR_BM25_CW 16
#endif
constexpr int kConsumerWarps = THR_BM25_CW;
constexpr int kConsumers = kConsumerWarps * 32;                    // 512
constexpr int kThreads = kConsumers + 32;                          // + producer warp
constexpr int kMaxBlkDocs = 16384;
constexpr int kRingCap = 12288;                  // postings in the ring (96 KB)
constexpr int kStepCap = 4096;                   // postings per step
constexpr int kSlots = 8;                        // steps in flight
constexpr int kMaxSelB = 256;
constexpr int kCandCap = kStepCap + 2 * kMaxSelB;  // a single-step range can append a whole step after a compaction
constexpr int kMaxSeg = kMaxTerms;               // one segment per term and step
constexpr int kPtrDepth = 12;                    // skip entries prefetched ahead
constexpr int kPtrRing = 16;

struct Posting { uint32_t doc; float imp; };

// A work unit: one query restricted to the doc ranges [r0, r1).
struct Unit { int q; int r0; int r1; unsigned cost; };
constexpr int kMaxUnitsPerQuery = 16;

enum { kLast = 1, kEou = 2, kSingle = 4 };

struct StepMeta {
  int nseg;
  int flags;             // kLast: last step of its range; kSingle: the only step of its range; kEou: end of unit
  int range;             // range index
  int add_bound;         // upper bound of docs the range can append (valid on the last step)
  uint2 seg[kMaxSeg];    // .x = first valid posting (ring index) | count << 16,  .y = idf of the term (float bits)
};
static_assert(kRingCap <= 65536 && kStepCap < 65536, "segment start / count are packed into 16 bits each");
static_assert(sizeof(StepMeta) % 8 == 0, "mbarriers follow the metadata and need 8-byte alignment");

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
  thr_dev_status* status;
};

__device__ __forceinline__ void bar_consumers() {
  asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
}
// Barrier over the consumer threads that also ORs a predicate: every thread gets the same answer.
__device__ __forceinline__ bool bar_consumers_or(bool p) {
  uint32_t r;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 p, %1, 0;\n\t"
      "bar.red.or.pred q, 1, %2, p;\n\t"
      "selp.u32 %0, 1, 0, q;\n\t}"
      : "=r"(r)
      : "r"((uint32_t)p), "n"(kConsumers)
      : "memory");
  return r != 0;
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      :
      : "r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
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
  constexpr int kPer = (kCandCap + kConsumers - 1) / kConsumers;
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

constexpr size_t kSmemRing = (size_t)kRingCap * sizeof(Posting);
constexpr size_t kSmemAcc = (size_t)kMaxBlkDocs * 4;
constexpr size_t kSmemCand = (size_t)kCandCap * 8;
constexpr size_t kSmemMeta = (size_t)kSlots * sizeof(StepMeta);
constexpr size_t kSmemBars = (size_t)2 * kSlots * 8;
constexpr size_t kSmemPtr = (size_t)kPtrRing * 32 * 8;
constexpr size_t kSmemMisc = 256 * 4 + kMaxTerms * 8 + 8 + 8 * 4;
constexpr size_t kBm25Smem = kSmemRing + kSmemAcc + kSmemCand + kSmemMeta + kSmemBars + kSmemPtr + kSmemMisc + 256;

__global__ void __launch_bounds__(kThreads, 1) bm25_kernel(const Bm25Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // carve-up (every region keeps 8-byte alignment)
  Posting* ring = (Posting*)gen;
  float* acc = (float*)(gen + kSmemRing);
  uint64_t* cand = (uint64_t*)(acc + kMaxBlkDocs);
  StepMeta* meta = (StepMeta*)(cand + kCandCap);
  uint64_t* bars = (uint64_t*)(meta + kSlots);                                // full, empty
  long long* pring = (long long*)(bars + 2 * kSlots);                         // [kPtrRing][32]
  uint32_t* hist = (uint32_t*)(pring + kPtrRing * 32);                        // 256
  int* q_term = (int*)(hist + 256);                                           // kMaxTerms
  float* q_idf = (float*)(q_term + kMaxTerms);                                // kMaxTerms
  unsigned long long* s_prefix = (unsigned long long*)(q_idf + kMaxTerms);
  int* s_int = (int*)(s_prefix + 1);  // [0]=want [1]=cnt(compact) [2]=cand count [3]=unit [4]=nterms

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto full_bar = [&](int s) { return smem_u32(&bars[s]); };
  auto empty_bar = [&](int s) { return smem_u32(&bars[kSlots + s]); };

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kConsumerWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < kMaxBlkDocs; i += kThreads) acc[i] = 0.f;
  __syncthreads();

  uint32_t it = 0;  // step counter, advances identically in all three roles
  const int R = a.blk_docs;
  // producer-only ring state
  int head = 0;          // next free posting in the ring
  uint32_t oldest = 0;   // oldest step whose data may still be in use
  int my_begin = 0;      // lane j < kSlots: ring offset of the step in slot j

  for (;;) {
    // ---- fetch the next unit (whole CTA) ----
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
      const float my_idf = lane < nterms ? q_idf[lane] : 0.f;
      const int64_t* row = a.skip + (size_t)(my_term < 0 ? 0 : my_term) * a.n_blk + r_begin;
      const int n_ranges = r_end - r_begin;
      const uint32_t my_pr = smem_u32(pring + lane);
      auto issue_ptr = [&](int e) {  // skip entry e of this unit: row[e], e in [0, n_ranges]
        if (my_term >= 0 && e <= n_ranges) cp_async_8(my_pr + (uint32_t)(e % kPtrRing) * 256u, row + e);
        cp_async_commit();
      };
      auto wait_step = [&](uint32_t j) {  // until the consumers released step j
        mbar_wait_relaxed(empty_bar(j % kSlots), (j / kSlots) & 1u, a.status, 400);
      };
      // Room for n postings (even) in the ring for step `it`; returns the ring offset.
      auto alloc = [&](int n) -> int {
        while (oldest + kSlots <= it) { wait_step(oldest); ++oldest; }  // the slot itself must be free
        for (;;) {
          int pos = -1;
          if (oldest == it) { head = 0; pos = 0; }  // nothing outstanding
          else {
            const int tb = __shfl_sync(0xffffffffu, my_begin, (int)(oldest % kSlots));
            if (head >= tb) {            // occupied: [tb, head)
              if (n <= kRingCap - head) pos = head;
              else if (n < tb) pos = 0;
            } else if (n < tb - head) {  // occupied: [tb, cap) + [0, head)
              pos = head;
            }
          }
          if (pos >= 0) {
            head = pos + n;
            if (lane == (int)(it % kSlots)) my_begin = pos;
            return pos;
          }
          wait_step(oldest);
          ++oldest;
        }
      };

      for (int e = 0; e <= kPtrDepth; ++e) issue_ptr(e);
      for (int i = 0; i < n_ranges; ++i) {
        issue_ptr(i + kPtrDepth + 1);
        cp_async_wait<kPtrDepth>();  // entries 0 .. i+1 have landed
        int64_t lo = 0, hi = 0;
        if (my_term >= 0) {
          lo = pring[(i % kPtrRing) * 32 + lane];
          hi = pring[((i + 1) % kPtrRing) * 32 + lane];
        }
        int64_t remaining = hi - lo;
        long long tot = remaining;
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, s);
        if (tot == 0) continue;  // nothing in this range for this query
        const int add_bound = (int)min((long long)R, tot);
        bool first_step = true;
        // Pack the terms' segments into steps of <= kStepCap postings, in term order (normally one step).
        for (;;) {
          const bool has = remaining > 0;
          const int slack = (int)(lo & 1);  // copies start at an even posting (16-byte granules)
          const long long need_ll = has ? ((slack + remaining + 1) & ~1LL) : 0;
          const int need = (int)min(need_ll, (long long)(kStepCap + 2));
          int incl = need;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += v;
          }
          const int before = incl - need;
          int take = 0, cp = 0;
          if (has) {
            if (incl <= kStepCap) { take = (int)remaining; cp = need; }
            else if (before < kStepCap) {  // first term that does not fit: take a piece if it is worth a copy
              const int room = kStepCap - before - slack - 1;
              if (room >= 64) { take = (int)min((long long)room, (long long)remaining); cp = (slack + take + 1) & ~1; }
            }
          }
          int cincl = cp;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, cincl, d);
            if (lane >= d) cincl += v;
          }
          const int total_cp = __shfl_sync(0xffffffffu, cincl, 31);
          const int off = cincl - cp;
          const unsigned have = __ballot_sync(0xffffffffu, take > 0);
          const unsigned still = __ballot_sync(0xffffffffu, remaining - take > 0);
          const int s = it % kSlots;
          const int pos = alloc(total_cp);
          if (take > 0) {
            const int g = __popc(have & ((1u << lane) - 1));
            meta[s].seg[g] = make_uint2((uint32_t)(pos + off + slack) | ((uint32_t)take << 16), __float_as_uint(my_idf));
          }
          __syncwarp();
          if (lane == 0) {
            meta[s].nseg = __popc(have);
            meta[s].flags = (still ? 0 : kLast) | ((first_step && !still) ? kSingle : 0);
            meta[s].range = r_begin + i;
            meta[s].add_bound = add_bound;
            // metadata is written with generic stores; the arrive has release semantics
            mbar_arrive_expect_tx(full_bar(s), (uint32_t)total_cp * 8u);
          }
          __syncwarp();
          if (take > 0)
            bulk_g2s(smem_u32(ring + pos + off), a.post + (lo - slack), (uint32_t)cp * 8u, full_bar(s));
          lo += take;
          remaining -= take;
          first_step = false;
          ++it;
          if (!still) break;
        }
      }
      // end-of-unit marker
      {
        const int s = it % kSlots;
        (void)alloc(0);
        if (lane == 0) {
          meta[s].nseg = 0;
          meta[s].flags = kEou;
          mbar_arrive(full_bar(s));
        }
        ++it;
        __syncwarp();
      }
      cp_async_wait<0>();
    } else {
      // ======================= consumers =======================
      // The hot loops address shared memory with 32-bit shared-space addresses (ld/st.shared).
      float tau = 0.f;  // only score > 0 is eligible; raised by compactions
      volatile int* v_cnt = &s_int[2];
      const uint32_t ring_addr = smem_u32(ring), acc_addr = smem_u32(acc);
      // rare path: one candidate per calling lane
      auto emit1 = [&](float v, uint32_t doc) {
        const int pos = atomicAdd(&s_int[2], 1);
        if (pos < kCandCap) cand[pos] = pack_key(v, doc);
        else dev_report(a.status, THR_EOVERFLOW, 430, pos);  // excluded by the capacity votes; keep it loud
      };
      auto compact = [&]() {
        const uint64_t T = block_compact_topk(cand, *v_cnt, a.k, hist, s_prefix, &s_int[0], &s_int[1], tid);
        tau = key_score(T);
        if (tid == 0) s_int[2] = s_int[1];
        bar_consumers();
      };
      for (;;) {
        const int s = it % kSlots;
        const uint32_t ph = (it / kSlots) & 1u;
        mbar_wait(full_bar(s), ph, a.status, 410);
        ++it;
        const uint32_t meta_addr = smem_u32(&meta[s]);
        int nseg, flags, range, add_bound;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(nseg), "=r"(flags), "=r"(range), "=r"(add_bound) : "r"(meta_addr));
        if (flags & kEou) {
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(s));
          break;
        }
        const uint32_t doc0 = (uint32_t)range << a.blk_shift;
        const uint32_t acc0 = acc_addr - doc0 * 4u;  // &acc[doc - doc0] == acc0 + doc * 4 (mod 2^32)
        const bool single = (flags & kSingle) != 0;
        const bool last = (flags & kLast) != 0;
        // tight: even a list compacted to k could overflow (only ranges spread over several steps) — those
        // vote per scan chunk instead.  Otherwise a true vote implies count > k, which the compaction needs.
        const bool tight = a.k + add_bound > kCandCap;
        bool vote = false;
        // ---- pass 1: accumulate, one term at a time ----
        uint32_t seg_addr = meta_addr + 16u;
        for (int g = 0; g < nseg; ++g, seg_addr += 8u) {
          uint32_t pw;
          float w;
          asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(pw), "=f"(w) : "r"(seg_addr));
          const uint32_t end = ring_addr + ((pw & 0xffffu) + (pw >> 16)) * 8u;
          for (uint32_t pa = ring_addr + ((pw & 0xffffu) + (uint32_t)tid) * 8u; pa < end; pa += kConsumers * 8u) {
            uint32_t doc;
            float imp, old;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(doc), "=f"(imp) : "r"(pa));
            const uint32_t aa = acc0 + doc * 4u;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(old) : "r"(aa));
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(aa), "f"(__fadd_rn(old, __fmul_rn(w, imp))) : "memory");
            // the first posting to touch a doc owns it in pass 2
            if (single && old == 0.f) asm volatile("st.shared.b32 [%0], %1;" ::"r"(pa), "r"(doc | 0x80000000u) : "memory");
          }
          // term boundary: the next term (or the next step) may hit the same docs
          if (g + 1 == nseg && last && !tight) vote = bar_consumers_or(*v_cnt + add_bound > kCandCap);
          else bar_consumers();
        }
        if (!single) {
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(s));  // this warp no longer reads the step's postings
        }
        if (!last) continue;

        // ---- range complete: collect its docs ----
        if (vote) compact();
        if (single) {
          // sparse: owners read the final score, append if > tau, re-zero the slot
          seg_addr = meta_addr + 16u;
          for (int g = 0; g < nseg; ++g, seg_addr += 8u) {
            uint32_t pw;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(pw) : "r"(seg_addr));
            const uint32_t end = ring_addr + ((pw & 0xffffu) + (pw >> 16)) * 8u;
            for (uint32_t pa = ring_addr + ((pw & 0xffffu) + (uint32_t)tid) * 8u; pa < end; pa += kConsumers * 8u) {
              uint32_t d;
              asm volatile("ld.shared.b32 %0, [%1];" : "=r"(d) : "r"(pa));
              if (d & 0x80000000u) {
                const uint32_t doc = d & 0x7fffffffu;
                const uint32_t aa = acc0 + doc * 4u;
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(aa));
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(aa), "f"(0.f) : "memory");
                if (v > tau) emit1(v, doc);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(empty_bar(s));
        } else {
          // the range came in several steps: scan the accumulators
          for (int c0 = 0; c0 < R; c0 += kConsumers * 4) {  // uniform trip count: the vote is a block barrier
            if (tight && bar_consumers_or(*v_cnt + kConsumers * 4 > kCandCap)) compact();
            const int c = c0 + tid * 4;
            if (c < R) {
              float4* p4 = reinterpret_cast<float4*>(&acc[c]);
              const float4 v = *p4;
              if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
                *p4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const uint32_t d = doc0 + (uint32_t)c;
                if (v.x > tau) emit1(v.x, d);
                if (v.y > tau) emit1(v.y, d + 1);
                if (v.z > tau) emit1(v.z, d + 2);
                if (v.w > tau) emit1(v.w, d + 3);
              }
            }
          }
        }
        bar_consumers();  // slots re-zeroed and appends visible before the next range accumulates
      }

      // ---- end of unit: final top-k, sorted ----
      bar_consumers();
      int n = *v_cnt;
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

__global__ void bm25_df_kernel(const int64_t* skip, int n_blk, int V, int64_t* df) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  df[t] = skip[(size_t)(t + 1) * n_blk] - skip[(size_t)t * n_blk];
}

// cost[q] = total postings of the query's terms.
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

// Single block: cut queries into units of roughly equal cost.  A range costs its postings plus a fixed
// per-range overhead (kRangeCost postings' worth of pipeline work), so light queries are split as well.
constexpr unsigned long long kRangeCost = 512;
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
  for (int q = tid; q < B; q += 1024) part += (keys[q] >> 32) + kRangeCost * (unsigned long long)n_blk;
  atomicAdd(&s_tot, part);
  __syncthreads();
  unsigned long long target = s_tot / (unsigned long long)(num_sms * 3) + 1;
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

}  // namespace

struct thr_bm25_state {
  const int64_t* skip;
  const void* post;
  const float* idf;
  int64_t n_docs;
  int n_blk, blk_docs, blk_shift, V;
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

int thr_bm25_index_set(thr_handle* h, const int64_t* skip, const void* postings, const float* idf,
                       int64_t n_docs, int32_t n_blk, int32_t blk_docs, int32_t V, int64_t id_base) {
  if (!h) return THR_EINVAL;
  cudaSetDevice(h->device);
  THR_REQUIRE(h, skip && postings && idf, "thr_bm25_index_set: NULL argument");
  THR_REQUIRE(h, n_docs >= 1 && V >= 1 && n_blk >= 1, "thr_bm25_index_set: empty index");
  int shift = 0;
  while ((1 << shift) < blk_docs) ++shift;
  if ((1 << shift) != blk_docs || blk_docs < 1024 || blk_docs > kMaxBlkDocs)
    return thr_fail(h, THR_EUNSUPPORTED, "thr_bm25_index_set: blk_docs = %d must be a power of two in [1024, %d]",
                    blk_docs, kMaxBlkDocs);
  THR_REQUIRE(h, (int64_t)n_blk * blk_docs >= n_docs && (int64_t)(n_blk - 1) * blk_docs < n_docs,
              "thr_bm25_index_set: n_blk does not match n_docs / blk_docs");
  THR_REQUIRE(h, n_docs < ((int64_t)1 << 31), "thr_bm25_index_set: more than 2^31 docs per shard (bit 31 of a staged doc id is the owner flag)");
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
  a.skip = st->skip; a.post = (const Posting*)st->post; a.idf = st->idf; a.n_docs = st->n_docs;
  a.n_blk = st->n_blk; a.blk_docs = st->blk_docs; a.blk_shift = st->blk_shift; a.V = st->V;
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
