// common.cuh — shared host/device helpers of libmsgpu (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>
#include <utility>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

namespace msgpu {

constexpr int kNumSMs = 148; // B200: 2 dies x 74 SMs; grids are sized in multiples of this (persistent / grid-stride)

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define MSGPU_CUDA(expr)                                                                                    \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      throw ::msgpu::CudaError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                               ":" + std::to_string(__LINE__) + ")");                                       \
    }                                                                                                       \
  } while (0)

#define MSGPU_LAUNCH_CHECK() MSGPU_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------------------------------------
// Device arena: one growing list of cudaMalloc'd slabs, bump-allocated, reset per build.  Slabs are kept across builds
// (HBM is 180 GB; re-running the same workload must not pay cudaMalloc again) and coalesced into one slab of the
// high-water size on reset.
// ---------------------------------------------------------------------------------------------------------------------
class Arena {
public:
  ~Arena() { release(); }
  void release() {
    for (auto &s : m_slabs) cudaFree(s.base);
    m_slabs.clear();
    m_used = m_high = 0;
    if (m_counter) cudaFree(m_counter);
    m_counter = nullptr;
    if (m_host_scratch) cudaFreeHost(m_host_scratch);
    m_host_scratch = nullptr;
    for (int k = 0; k < 4; ++k) {
      if (m_scan_state[k]) cudaFree(m_scan_state[k]);
      m_scan_state[k]       = nullptr;
      m_scan_state_bytes[k] = 0;
    }
    if (m_side) cudaStreamDestroy(m_side);
    m_side = nullptr;
    for (auto ev : m_events) cudaEventDestroy(ev);
    m_events.clear();
  }
  // A second, high-priority stream of the context for work that runs BESIDE the build's stream (the emission kernel of
  // one chunk underneath the tile kernel of the next, pair_fused.cu), and events to order the two.  The caller joins the
  // side stream back into the build's stream before the build returns.
  cudaStream_t side_stream() {
    if (!m_side) {
      int least = 0, greatest = 0;
      MSGPU_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      MSGPU_CUDA(cudaStreamCreateWithPriority(&m_side, cudaStreamNonBlocking, greatest));
    }
    return m_side;
  }
  cudaEvent_t sync_event(size_t i) {
    while (m_events.size() <= i) {
      cudaEvent_t ev;
      MSGPU_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      m_events.push_back(ev);
    }
    return m_events[i];
  }
  // Status words of the single-pass scans (scan.cuh): regions that outlive reset() and that nothing else writes, one
  // per slot (up to four scans run concurrently in one launch).  Every call hands out the next epoch; flags written
  // under an older epoch read as "not ready", so a region is cleared only when it grows or when the 16-bit epoch wraps.
  // Tickets: a scan of `n_tickets` tiles numbers them in the order its blocks start, from ONE counter per slot that
  // only ever grows; the host knows how many tickets every earlier scan of the slot took (exactly one per tile), so
  // a scan subtracts that base and nothing ever has to be cleared — a scan that is prepared but never launched, or
  // one with a single tile, takes no ticket and leaves the bookkeeping untouched (unsigned wrap-around is harmless).
  void *scan_state(int slot, size_t bytes, cudaStream_t st, unsigned *epoch, unsigned n_tickets, unsigned **ticket,
                   unsigned *ticket_base) {
    *ticket      = counter() + 8 + slot;
    *ticket_base = m_ticket_base[slot];
    m_ticket_base[slot] += n_tickets;
    if (bytes > m_scan_state_bytes[slot]) {
      if (m_scan_state[slot]) {
        MSGPU_CUDA(cudaStreamSynchronize(st));
        cudaFree(m_scan_state[slot]);
      }
      m_scan_state_bytes[slot] = bytes < (size_t(1) << 18) ? (size_t(1) << 18) : bytes * 2;
      MSGPU_CUDA(cudaMalloc(&m_scan_state[slot], m_scan_state_bytes[slot]));
      MSGPU_CUDA(cudaMemsetAsync(m_scan_state[slot], 0, m_scan_state_bytes[slot], st));
    }
    if (++m_epoch >= (1u << 16)) {
      for (int k = 0; k < 4; ++k)
        if (m_scan_state[k]) MSGPU_CUDA(cudaMemsetAsync(m_scan_state[k], 0, m_scan_state_bytes[k], st));
      m_epoch = 1;
    }
    *epoch = m_epoch;
    return m_scan_state[slot];
  }
  static constexpr size_t kHostScratch = 16384;
  // a small pinned host buffer for the few words a stage reads back (a pageable destination makes the copy take the
  // driver's staging path); valid until the next call
  void *host_scratch() {
    if (!m_host_scratch) MSGPU_CUDA(cudaMallocHost(&m_host_scratch, kHostScratch));
    return m_host_scratch;
  }
  // zero-initialised device counters that outlive reset(): kernels that use them leave them at zero again
  unsigned *counter() {
    if (!m_counter) {
      MSGPU_CUDA(cudaMalloc(&m_counter, 512)); // [0..7] kernels' own counters, [8..11] ticket counters of the scans
      MSGPU_CUDA(cudaMemset(m_counter, 0, 512));
    }
    return m_counter;
  }
  // start a new build: if the previous one spilled into several slabs, replace them by one of the total size
  void reset() {
    if (m_slabs.size() > 1) {
      size_t total = 0;
      for (auto &s : m_slabs) total += s.size;
      for (auto &s : m_slabs) cudaFree(s.base);
      m_slabs.clear();
      Slab  s{nullptr, total, 0};
      MSGPU_CUDA(cudaMalloc(&s.base, total));
      m_slabs.push_back(s);
    }
    for (auto &s : m_slabs) s.used = 0;
    m_used = 0;
  }
  template <class T> T *alloc(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    for (auto &s : m_slabs) {
      if (s.size - s.used >= bytes) {
        void *p = static_cast<char *>(s.base) + s.used;
        s.used += bytes;
        m_used += bytes;
        if (m_used > m_high) m_high = m_used;
        return static_cast<T *>(p);
      }
    }
    size_t want = bytes > (size_t(64) << 20) ? bytes : (size_t(64) << 20);
    if (want < m_high / 2) want = m_high / 2;
    Slab s{nullptr, want, 0};
    MSGPU_CUDA(cudaMalloc(&s.base, want));
    m_slabs.push_back(s);
    return alloc<T>(n);
  }
  size_t high_water() const { return m_high; }
  // Stack discipline for temporaries: everything allocated after mark() is given back by release_to() — the memory
  // is reused by later allocations of the same build.  Stream order makes this safe without synchronisation: the
  // kernels that used the temporaries were enqueued before the kernels that will reuse their memory.
  struct Mark {
    size_t n_slabs, used_last, used_total;
  };
  Mark mark() const { return Mark{m_slabs.size(), m_slabs.empty() ? 0 : m_slabs.back().used, m_used}; }
  void release_to(const Mark &m) {
    if (m.n_slabs == 0) {
      for (auto &s : m_slabs) s.used = 0;
    } else {
      for (size_t i = m.n_slabs; i < m_slabs.size(); ++i) m_slabs[i].used = 0;
      // (alloc() fills the first slab that has room, so earlier slabs may have grown too: only the slab that was last
      //  at mark time can be rolled back exactly; allocations that landed in older slabs stay until reset())
      if (m_slabs[m.n_slabs - 1].used > m.used_last) m_slabs[m.n_slabs - 1].used = m.used_last;
    }
    m_used = m.used_total;
  }

private:
  struct Slab {
    void  *base;
    size_t size, used;
  };
  std::vector<Slab> m_slabs;
  size_t            m_used = 0, m_high = 0;
  unsigned         *m_counter = nullptr;
  void             *m_host_scratch = nullptr;
  void             *m_scan_state[4] = {nullptr, nullptr, nullptr, nullptr};
  size_t            m_scan_state_bytes[4] = {0, 0, 0, 0};
  unsigned          m_epoch = 0;
  unsigned          m_ticket_base[4] = {0, 0, 0, 0};
  cudaStream_t             m_side = nullptr;
  std::vector<cudaEvent_t> m_events;
};

// ---------------------------------------------------------------------------------------------------------------------
// Launch accounting + optional per-kernel timing (CUDA events on the launching stream).  Every kernel launch of the
// library goes through MSGPU_LAUNCH, so `launches` is the library's own count of what it put on the GPU; when
// `enabled`, each launch is bracketed by two events and bench.py reads per-kernel durations for the roofline.
// ---------------------------------------------------------------------------------------------------------------------
struct Profiler {
  struct Rec {
    const char *name;
    cudaEvent_t a, b;
  };
  bool                     enabled  = false;
  uint64_t                 launches = 0;
  std::vector<Rec>         recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e;
    MSGPU_CUDA(cudaEventCreate(&e));
    return e;
  }
  void clear() {
    for (auto &r : recs) {
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.clear();
  }
  ~Profiler() {
    clear();
    for (auto e : pool) cudaEventDestroy(e);
  }
};

inline Profiler *&current_profiler() {
  static thread_local Profiler *p = nullptr;
  return p;
}

struct KernelScope {
  Profiler    *prof;
  cudaStream_t st;
  cudaEvent_t  stop = nullptr;
  KernelScope(const char *name, cudaStream_t s) : prof(current_profiler()), st(s) {
    if (!prof) return;
    ++prof->launches;
    if (prof->enabled) {
      Profiler::Rec r{name, prof->get_event(), prof->get_event()};
      stop = r.b;
      cudaEventRecord(r.a, st);
      prof->recs.push_back(r);
    }
  }
  ~KernelScope() {
    if (stop) cudaEventRecord(stop, st);
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Programmatic dependent launches.  A construction is a chain of ~30 short kernels, each depending on the one before
// it; between two of them the GPU sits idle for the launch latency of the second.  Every kernel of the library opens
// with pdl_sync(): wait until the grid before it in the stream has COMPLETED and its writes are visible
// (griddepcontrol.wait — a no-op for a kernel launched the ordinary way), then allow the grid behind it to be scheduled
// (griddepcontrol.launch_dependents).  launch_k() launches with the programmatic-stream-serialization attribute, so
// the CTAs of the next kernel are placed on the SMs while this one drains and start the moment it has finished,
// instead of being launched only then.  Ordering and visibility are exactly those of an ordinary stream (every thread
// waits before it touches anything); copies, memsets and events between kernels serialise as always.  MSGPU_PDL=0
// launches without the attribute.
// ---------------------------------------------------------------------------------------------------------------------
inline bool pdl_enabled() {
  static const bool on = [] {
    const char *e = std::getenv("MSGPU_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <class... KArgs, class... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim          = grid;
  cfg.blockDim         = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream           = st;
  cudaLaunchAttribute attr[1];
  attr[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs    = attr;
  cfg.numAttrs = pdl_enabled() ? 1u : 0u;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...); // (MSGPU_LAUNCH checks the error)
}
#endif

#define MSGPU_LAUNCH(name, stream, ...)        \
  do {                                         \
    ::msgpu::KernelScope _ks(name, stream);    \
    __VA_ARGS__;                               \
    MSGPU_LAUNCH_CHECK();                      \
  } while (0)

inline int grid_for(size_t n_items, int block, int max_waves = 8) {
  // grid = multiple of the SM count, capped at `max_waves` x (SMs x resident blocks); kernels are grid-stride
  size_t need   = (n_items + block - 1) / block;
  size_t per_sm = 2048 / block;
  size_t cap    = size_t(kNumSMs) * per_sm * max_waves;
  if (need > cap) need = cap;
  size_t g = ((need + kNumSMs - 1) / kNumSMs) * kNumSMs;
  return static_cast<int>(g ? g : kNumSMs);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned warp_id() { return threadIdx.x >> 5; }

template <class T> __device__ __forceinline__ T warp_inclusive_scan(T v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane_id() >= (unsigned)d) v += n;
  }
  return v;
}

template <class T> __device__ __forceinline__ T warp_reduce_sum(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32).  `smem` holds 33 T.
// Returns the exclusive prefix of the calling thread; *total receives the block sum.  Ends with a __syncthreads()
// so `smem` can be reused immediately.
template <class T> __device__ __forceinline__ T block_exclusive_scan(T v, T *smem, T *total) {
  T        incl = warp_inclusive_scan(v);
  unsigned w = warp_id(), l = lane_id(), nw = blockDim.x >> 5;
  if (l == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    T x  = l < nw ? smem[l] : T(0);
    T xs = warp_inclusive_scan(x);
    if (l < nw) smem[l] = xs - x;
    if (l == nw - 1) smem[32] = xs;
  }
  __syncthreads();
  T res  = smem[w] + incl - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

// 128-bit streaming loads / stores: the inputs of the big passes are read once, so keep them out of L1
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// One 32-byte record (a PairRow / VmRow: one L2 sector) fetched by ONE 256-bit read-only load (sm_100: SASS
// LDG.E.ENL2.256.CONSTANT) instead of two 128-bit ones.  The gathers of the emission and chaining kernels hit a
// different line per lane, so every load instruction costs one L1 wavefront per lane: one instruction per record
// halves them (MSGPU_LD256=0 keeps the two-load form for the A/B).  `p` must be 32-byte aligned.
#ifndef MSGPU_LD256
#define MSGPU_LD256 1
#endif
__device__ __forceinline__ void ld_row32(const void *p, uint4 &x, uint4 &y) {
#if MSGPU_LD256 == 2 // development A/B: the same load without allocating the line in L1
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w), "=r"(y.x), "=r"(y.y), "=r"(y.z), "=r"(y.w)
               : "l"(p));
#elif MSGPU_LD256
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w), "=r"(y.x), "=r"(y.y), "=r"(y.z), "=r"(y.w)
               : "l"(p));
#else
  const uint4 *q = reinterpret_cast<const uint4 *>(p);
  x = __ldg(q);
  y = __ldg(q + 1);
#endif
}

// 64-bit mix (splitmix64 finaliser) — used to spread keys over the open-addressing tables
__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
#endif // __CUDACC__

} // namespace msgpu
