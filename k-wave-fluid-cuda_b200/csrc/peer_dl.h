// Peer-memory all-to-all of slab-decomposed runs: copy-engine pushes into IPC-mapped peer buffers over NVLink / NVSwitch,
// ordered by stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) on flags in a POSIX shared-memory
// segment that every rank of the node maps and registers with CUDA.  No SM is used for the exchange, so it overlaps the
// compute kernels without competing for CTA slots (an NCCL send/recv kernel can only start at a kernel boundary of the
// persistent compute grids, and then takes SMs away from them).  One node only (one process per GPU).
//
// The driver entry points are resolved at run time from libcuda.so.1 (the library links against the runtime only).
#pragma once
#include <cuda.h>  // driver API TYPES only (CUstreamBatchMemOpParams); entry points are resolved with dlsym
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>

namespace kw {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerBufs = 8;  // S[0..3], R[0..3]

struct DrvApi {
  // CUresult cuStreamWaitValue32(CUstream, CUdeviceptr, cuuint32_t value, unsigned flags)
  int (*WaitValue32)(cudaStream_t, unsigned long long, uint32_t, unsigned) = nullptr;
  int (*WriteValue32)(cudaStream_t, unsigned long long, uint32_t, unsigned) = nullptr;
  // CUresult cuStreamBatchMemOp(CUstream, unsigned count, CUstreamBatchMemOpParams*, unsigned flags): one driver call for the
  // P-1 waits or writes of an exchange (the host enqueues ~40 operations per exchange at 8 ranks otherwise)
  int (*BatchMemOp)(cudaStream_t, unsigned, CUstreamBatchMemOpParams*, unsigned) = nullptr;
  bool ok = false;
  std::string error;
};
enum { kWaitGeq = 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */, kWriteDefault = 0x0 };

inline DrvApi& drv_api() {
  static DrvApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    api.error = "libcuda.so.1 not found";
    return api;
  }
  const char* wait_names[] = {"cuStreamWaitValue32_v2", "cuStreamWaitValue32"};
  const char* write_names[] = {"cuStreamWriteValue32_v2", "cuStreamWriteValue32"};
  for (const char* n : wait_names)
    if (!api.WaitValue32) api.WaitValue32 = reinterpret_cast<decltype(api.WaitValue32)>(dlsym(h, n));
  for (const char* n : write_names)
    if (!api.WriteValue32) api.WriteValue32 = reinterpret_cast<decltype(api.WriteValue32)>(dlsym(h, n));
  const char* batch_names[] = {"cuStreamBatchMemOp_v2", "cuStreamBatchMemOp"};
  for (const char* n : batch_names)
    if (!api.BatchMemOp) api.BatchMemOp = reinterpret_cast<decltype(api.BatchMemOp)>(dlsym(h, n));
  if (getenv("KW_BATCH_MEMOPS") && atoi(getenv("KW_BATCH_MEMOPS")) == 0) api.BatchMemOp = nullptr;
  if (!api.WaitValue32 || !api.WriteValue32) {
    api.error = "cuStreamWaitValue32 / cuStreamWriteValue32 missing from the driver";
    return api;
  }
  api.ok = true;
  return api;
}

// what every rank of the run maps; plain integers only (lives in /dev/shm)
struct PeerShm {
  std::atomic<uint32_t> attached;                       // ranks that have written their handle
  std::atomic<uint32_t> opened;                         // ranks that have opened every peer handle (or failed)
  std::atomic<uint32_t> failed;                         // ranks whose set-up failed: everybody falls back to NCCL
  std::atomic<uint32_t> detached;                       // ranks that are done with the segment
  cudaIpcMemHandle_t handle[kPeerMaxRanks];             // arena of rank r
  // arrived[dst][buf][src] = n: the n-th use of buffer `buf` of rank `dst` holds the block of rank `src`
  volatile uint32_t arrived[kPeerMaxRanks][kPeerBufs][kPeerMaxRanks];
  // credit[src][buf][dst] = n: rank `dst` no longer reads the n-th content of its buffer `buf` (src may push use n + 1)
  volatile uint32_t credit[kPeerMaxRanks][kPeerBufs][kPeerMaxRanks];
};

struct PeerLink {
  bool active = false;
  int rank = 0, nranks = 1;
  std::string shm_name;
  PeerShm* shm = nullptr;        // host mapping
  PeerShm* shm_dev = nullptr;    // device-side address of the same segment (stream memory operations)
  char* peer_base[kPeerMaxRanks] = {};  // arena of every rank as seen from this process (own arena at [rank])
  uint32_t use[kPeerBufs] = {};  // how often each buffer has been the destination of an exchange
  std::string error;

  static bool wait_count(std::atomic<uint32_t>& a, uint32_t want, double timeout_s) {
    const auto t0 = std::chrono::steady_clock::now();
    while (a.load(std::memory_order_acquire) < want) {
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
      std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    return true;
  }

  // `tag`: bytes shared by all ranks of this run and by nobody else (the ncclUniqueId); arena = this rank's cudaMalloc'ed
  // block holding the kPeerBufs exchange buffers.  Collective over the ranks; returns false (with `error`) when the peer
  // path is unavailable -- the caller then keeps the NCCL exchange.  ALL RANKS TAKE THE SAME DECISION: `agree(ok)` is a collective
  // of the caller (an all-reduce with min over the run's communicator) that returns true only when every rank passed true; it is
  // called exactly twice by every rank, whatever happened locally -- a rank whose shm_open or registration failed no longer leaves
  // the others waiting for a counter, and nobody ends up on the peer path while a neighbour fell back to NCCL.  `want` carries the
  // per-process settings (KW_PEER) into the same agreement.
  template <class Agree> bool setup(const void* tag, size_t tag_bytes, int rank_, int nranks_, void* arena, bool want, Agree&& agree) {
    const bool ok1 = want && setup_local(tag, tag_bytes, rank_, nranks_, arena);
    if (!want && error.empty()) error = "disabled by KW_PEER=0";
    const bool all1 = agree(ok1);  // every rank has published its IPC handle (or nobody continues)
    if (shm) shm_unlink(shm_name.c_str());  // every rank has it mapped (or gave up): the name can go, nothing leaks after a crash
    bool ok2 = all1;
    if (all1) {
      peer_base[rank] = static_cast<char*>(arena);
      for (int q = 0; q < nranks && ok2; ++q) {
        if (q == rank) continue;
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, shm->handle[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok2 = false, error = "cudaIpcOpenMemHandle failed for rank " + std::to_string(q);
        }
        peer_base[q] = static_cast<char*>(p);
      }
    }
    const bool all2 = agree(ok2);
    if (!all2) {
      if (error.empty()) error = "another rank could not set up the peer path";
      if (shm) teardown();
      return false;
    }
    active = true;
    return true;
  }
  bool setup_local(const void* tag, size_t tag_bytes, int rank_, int nranks_, void* arena) {
    rank = rank_, nranks = nranks_;
    if (nranks > kPeerMaxRanks) return fail_local("more ranks than kPeerMaxRanks");
    uint64_t hsh = 1469598103934665603ull;
    for (size_t i = 0; i < tag_bytes; ++i) hsh = (hsh ^ static_cast<const unsigned char*>(tag)[i]) * 1099511628211ull;
    char name[64];
    snprintf(name, sizeof name, "/kwave_b200_%016llx", (unsigned long long)hsh);
    shm_name = name;
    const int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
    if (fd < 0) return fail_local("shm_open failed");
    if (ftruncate(fd, sizeof(PeerShm)) != 0) {
      close(fd);
      return fail_local("ftruncate failed");
    }
    void* m = mmap(nullptr, sizeof(PeerShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (m == MAP_FAILED) return fail_local("mmap failed");
    shm = static_cast<PeerShm*>(m);  // a fresh segment is zero filled
    bool ok = drv_api().ok;
    if (!ok) error = drv_api().error;
    if (ok && cudaHostRegister(shm, sizeof(PeerShm), cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess) {
      cudaGetLastError();
      ok = false, error = "cudaHostRegister of the flag segment failed";
    }
    void* dptr = nullptr;
    if (ok && cudaHostGetDevicePointer(&dptr, shm, 0) != cudaSuccess) ok = false, error = "cudaHostGetDevicePointer failed";
    shm_dev = static_cast<PeerShm*>(dptr);
    if (ok && cudaIpcGetMemHandle(&shm->handle[rank], arena) != cudaSuccess) {
      cudaGetLastError();
      ok = false, error = "cudaIpcGetMemHandle failed";
    }
    return ok;
  }
  bool fail_local(const char* why) {
    error = why;
    if (shm) shm->failed.fetch_add(1);
    return false;
  }
  void teardown() {
    for (int q = 0; q < nranks; ++q)
      if (q != rank && peer_base[q]) cudaIpcCloseMemHandle(peer_base[q]);
    if (shm) {
      cudaHostUnregister(shm);
      munmap(shm, sizeof(PeerShm));
      shm_unlink(shm_name.c_str());
    }
    shm = nullptr, active = false;
  }
  // `count` 32-bit waits (>= value) or writes on the flags flag(r), r != rank, in one driver call when available
  template <class F> int flag_ops(cudaStream_t stream, bool wait, uint32_t value, F&& flag) const {
    DrvApi& d = drv_api();
    if (d.BatchMemOp) {
      CUstreamBatchMemOpParams ops[kPeerMaxRanks];
      unsigned n = 0;
      for (int r = 0; r < nranks; ++r) {
        if (r == rank) continue;
        memset(&ops[n], 0, sizeof(ops[n]));
        if (wait) {
          ops[n].waitValue.operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
          ops[n].waitValue.address = (CUdeviceptr)dev_addr(flag(r));
          ops[n].waitValue.value = value;
          ops[n].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
        } else {
          ops[n].writeValue.operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
          ops[n].writeValue.address = (CUdeviceptr)dev_addr(flag(r));
          ops[n].writeValue.value = value;
          ops[n].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;
        }
        ++n;
      }
      return n ? d.BatchMemOp(stream, n, ops, 0) : 0;
    }
    for (int r = 0; r < nranks; ++r) {
      if (r == rank) continue;
      const int e = wait ? d.WaitValue32(stream, dev_addr(flag(r)), value, kWaitGeq) : d.WriteValue32(stream, dev_addr(flag(r)), value, kWriteDefault);
      if (e) return e;
    }
    return 0;
  }
  unsigned long long dev_addr(const volatile uint32_t* host_field) const {
    return reinterpret_cast<unsigned long long>(reinterpret_cast<const char*>(shm_dev) +
                                                (reinterpret_cast<const char*>(const_cast<const uint32_t*>(host_field)) - reinterpret_cast<const char*>(shm)));
  }
};

}  // namespace kw
