// vct_comm.cu -- the library-owned multi-GPU boundary (SURVEY.md 8b "pass entry points", 8e).
//
// The reference is single-GPU (one GL context, main.cpp:44); its loop `glClear -> Render -> glfwSwapBuffers`
// (main.cpp:77-94) is what a sharded frame replaces.  One process per GPU on one NVLink / NVSwitch node, or one
// process driving several devices (vct_create_multi).  Everything the exchange needs lives here, behind the C ABI:
//   * one SYMMETRIC SEGMENT per rank (cuMemCreate), mapped on every rank (cuMemMap of the peers' handles: NVLink
//     peer loads / stores) and bound to one MULTICAST object (cuMulticastCreate / cuMulticastBindMem: one
//     multimem.st lands in every rank's segment, replicated by the NVSwitch);
//   * the bootstrap that passes the allocation handles between processes (POSIX file descriptors over an abstract
//     unix socket, SCM_RIGHTS; rank 0 is the hub) -- no torch, no NCCL, no MPI;
//   * a DEVICE-SIDE BARRIER (one small kernel per rank: release-store an epoch into every peer's signal pad, then
//     acquire-spin on the own pad), enqueued in stream order, so no host thread ever waits for a peer;
//   * the segment layout:  [signal pads | voxel-exchange inbox (vct_voxelize.cu) | frame ring (rank 0 is the consumer)].
// The driver API is reached through cudaGetDriverEntryPoint, so the library keeps linking against the static
// runtime only and still loads on a machine without a driver (the CPU-side symbol tests).
#include <cuda.h>

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>

#include "vct_internal.h"

namespace vct {

// ------------------------------------------------------------------------------------------ driver entry points
struct Driver {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*DeviceGet)(CUdevice*, int);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*GetErrorString)(CUresult, const char**);
  bool ok = false;
};

static Driver g_drv;

static int load_driver(vct_context* c) {
  if (g_drv.ok) return VCT_OK;
  struct { const char* name; void** slot; } syms[] = {
      {"cuMemCreate", (void**)&g_drv.MemCreate}, {"cuMemRelease", (void**)&g_drv.MemRelease},
      {"cuMemAddressReserve", (void**)&g_drv.MemAddressReserve}, {"cuMemAddressFree", (void**)&g_drv.MemAddressFree},
      {"cuMemMap", (void**)&g_drv.MemMap}, {"cuMemUnmap", (void**)&g_drv.MemUnmap}, {"cuMemSetAccess", (void**)&g_drv.MemSetAccess},
      {"cuMemExportToShareableHandle", (void**)&g_drv.MemExportToShareableHandle},
      {"cuMemImportFromShareableHandle", (void**)&g_drv.MemImportFromShareableHandle},
      {"cuMemGetAllocationGranularity", (void**)&g_drv.MemGetAllocationGranularity},
      {"cuMulticastCreate", (void**)&g_drv.MulticastCreate}, {"cuMulticastAddDevice", (void**)&g_drv.MulticastAddDevice},
      {"cuMulticastBindMem", (void**)&g_drv.MulticastBindMem}, {"cuMulticastUnbind", (void**)&g_drv.MulticastUnbind},
      {"cuMulticastGetGranularity", (void**)&g_drv.MulticastGetGranularity}, {"cuDeviceGet", (void**)&g_drv.DeviceGet},
      {"cuDeviceGetAttribute", (void**)&g_drv.DeviceGetAttribute}, {"cuGetErrorString", (void**)&g_drv.GetErrorString}};
  for (auto& s : syms) {
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint(s.name, s.slot, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*s.slot)
      return set_error(c, VCT_ERR_CUDA, std::string("driver entry point not available: ") + s.name);
  }
  g_drv.ok = true;
  return VCT_OK;
}

static int check_cu(vct_context* c, CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return VCT_OK;
  const char* msg = nullptr;
  if (g_drv.GetErrorString) g_drv.GetErrorString(r, &msg);
  return set_error(c, VCT_ERR_CUDA, std::string(what) + ": " + (msg ? msg : "driver error ") + " (" + std::to_string((int)r) + ")");
}
#define VCT_CU(c, call)                                \
  do {                                                 \
    int _rc = check_cu((c), (call), #call);            \
    if (_rc) return _rc;                               \
  } while (0)

// ------------------------------------------------------------------------------------------ segment layout
constexpr size_t COMM_PADS = 4096;               // signal pads: [channel][16 ranks] x u32 at 256-byte channel stride; [2048] = fail flag
constexpr int COMM_CHANNELS = 4;

// one process, several devices: the members of a group share ONE multicast handle; the last one out releases it
struct InProcGroup { CUmemGenericAllocationHandle mc = 0; int refs = 0; };

struct Comm {
  InProcGroup* group = nullptr;
  int rank = 0, world = 1;
  bool multicast = false, in_process = false;
  size_t seg = 0;                                // bytes per rank segment (granularity aligned)
  size_t off_inbox = 0, inbox_bytes = 0, off_frames = 0, frame_bytes = 0, off_depth = 0, depth_bytes = 0;
  CUmemGenericAllocationHandle local = 0, mc = 0;
  std::vector<CUmemGenericAllocationHandle> peers;      // imported handles (index = rank; own slot = local)
  CUdeviceptr va = 0, mc_va = 0;
  CUdevice cu_dev = 0;
  std::vector<int> socks;                        // rank 0: one per peer (index = rank); others: [0] = hub
  uint32_t epoch[COMM_CHANNELS] = {0, 0, 0, 0};
  // sharded-frame ring (rank 0 consumes): host copies in flight
  unsigned long long frame_seq = 0;
  cudaEvent_t ev_band_done[3] = {nullptr, nullptr, nullptr}, ev_copied[3] = {nullptr, nullptr, nullptr};
  bool copy_pending[3] = {false, false, false};
  int V = 0, W = 0, H = 0; size_t exch_cap = 0; int shared_exchange = 0;   // the settings the segment was sized for
};

__host__ __device__ inline size_t pad_word(int channel, int rank) { return (size_t)channel * 64 + rank; }

// One thread per rank.  Thread t publishes this rank's arrival in rank t's pad, then waits for rank t's arrival
// in the own pad.  Epochs only grow, so the pads never need a reset.  The release / acquire pair at system scope
// orders everything this rank wrote before the barrier (previous kernels in the stream, including multimem.st and
// peer stores) before anything a peer reads after it.  A peer that never arrives costs `timeout_ns`, not a hung GPU.
__global__ void comm_barrier_kernel(uint32_t* __restrict__ va, size_t seg_words, int rank, int world, int channel,
                                    uint32_t epoch, unsigned long long timeout_ns) {
  const int t = threadIdx.x;
  if (t >= world) return;
  __threadfence_system();
  uint32_t* theirs = va + (size_t)t * seg_words + pad_word(channel, rank);
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t* mine = va + (size_t)rank * seg_words + pad_word(channel, t);
  unsigned long long t0 = 0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (unsigned spins = 0;; ++spins) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if ((spins & 1023u) == 1023u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) { va[(size_t)rank * seg_words + 512] = 1u + (uint32_t)t; break; }
    }
  }
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t inbox_bytes_for(const vct_context* c) {
  const size_t n = (size_t)c->P.V * c->P.V * c->P.V;
  size_t cap = c->exchange_cap_user ? c->exchange_cap_user : 32 * (size_t)c->P.V * c->P.V;
  if (cap > n) cap = n;
  if (c->shared_exchange == 1) return n * 16 + n / 8;          // dense symmetric accumulator + occupancy mask (multimem.red flavour)
  return 4096 + 2 * (size_t)c->shared_world * cap * 16;
}

// ------------------------------------------------------------------------------------------ bootstrap sockets
static int send_fds(int sock, const int* fds, int n) {
  char payload = 'F';
  struct iovec io = {&payload, 1};
  char ctrl[CMSG_SPACE(sizeof(int) * 32)];
  std::memset(ctrl, 0, sizeof(ctrl));
  struct msghdr msg = {};
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = CMSG_SPACE(sizeof(int) * n);
  struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int) * n);
  std::memcpy(CMSG_DATA(cm), fds, sizeof(int) * n);
  return sendmsg(sock, &msg, 0) == 1 ? 0 : -1;
}

static int recv_fds(int sock, int* fds, int n) {
  char payload = 0;
  struct iovec io = {&payload, 1};
  char ctrl[CMSG_SPACE(sizeof(int) * 32)];
  struct msghdr msg = {};
  msg.msg_iov = &io; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = CMSG_SPACE(sizeof(int) * n);
  if (recvmsg(sock, &msg, MSG_WAITALL) != 1) return -1;
  struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (!cm || cm->cmsg_type != SCM_RIGHTS || cm->cmsg_len != CMSG_LEN(sizeof(int) * n)) return -1;
  std::memcpy(fds, CMSG_DATA(cm), sizeof(int) * n);
  return 0;
}

static int send_all(int s, const void* p, size_t n) {
  const char* b = (const char*)p;
  while (n) { ssize_t k = send(s, b, n, MSG_NOSIGNAL); if (k <= 0) return -1; b += k; n -= (size_t)k; }
  return 0;
}
static int recv_all(int s, void* p, size_t n) {
  char* b = (char*)p;
  while (n) { ssize_t k = recv(s, b, n, 0); if (k <= 0) return -1; b += k; n -= (size_t)k; }
  return 0;
}

static socklen_t abstract_addr(struct sockaddr_un* a, const std::string& session) {
  std::memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  std::string name = "vct_b200_" + session;
  if (name.size() > sizeof(a->sun_path) - 2) name.resize(sizeof(a->sun_path) - 2);
  std::memcpy(a->sun_path + 1, name.data(), name.size());      // leading NUL: abstract namespace, nothing to unlink
  return (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + name.size());
}

static int bootstrap_connect(vct_context* c, Comm* m, const std::string& session, double timeout_s) {
  struct sockaddr_un addr;
  const socklen_t len = abstract_addr(&addr, session);
  struct timeval tv = {(time_t)timeout_s, 0};
  if (m->rank == 0) {
    int ls = socket(AF_UNIX, SOCK_STREAM, 0);
    if (ls < 0 || bind(ls, (struct sockaddr*)&addr, len) || listen(ls, 32)) {
      if (ls >= 0) close(ls);
      return set_error(c, VCT_ERR_STATE, std::string("vct_comm_init: cannot listen on session '") + session + "': " + strerror(errno));
    }
    setsockopt(ls, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    m->socks.assign(m->world, -1);
    for (int k = 1; k < m->world; ++k) {
      int s = accept(ls, nullptr, nullptr);
      int r = -1;
      if (s >= 0) setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
      if (s < 0 || recv_all(s, &r, 4) || r < 1 || r >= m->world || m->socks[r] >= 0) {
        if (s >= 0) close(s);
        close(ls);
        return set_error(c, VCT_ERR_STATE, "vct_comm_init: a peer did not join the session in time");
      }
      m->socks[r] = s;
    }
    close(ls);
  } else {
    int s = -1;
    struct timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
    while (true) {
      s = socket(AF_UNIX, SOCK_STREAM, 0);
      if (s >= 0 && connect(s, (struct sockaddr*)&addr, len) == 0) break;
      if (s >= 0) close(s);
      struct timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1);
      if ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec) > timeout_s)
        return set_error(c, VCT_ERR_STATE, "vct_comm_init: rank 0 is not listening on session '" + session + "'");
      usleep(20000);
    }
    setsockopt(s, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
    if (send_all(s, &m->rank, 4)) { close(s); return set_error(c, VCT_ERR_STATE, "vct_comm_init: hello failed"); }
    m->socks.assign(1, s);
  }
  return VCT_OK;
}

// host-side barrier over the bootstrap sockets (set-up only; frames use the device barrier).  `ok` is AND-ed.
static int host_barrier(vct_context* c, Comm* m, int ok, int* all_ok) {
  int v = ok, res = ok;
  if (m->in_process) { *all_ok = ok; return VCT_OK; }
  if (m->rank == 0) {
    for (int r = 1; r < m->world; ++r) { int x = 0; if (recv_all(m->socks[r], &x, 4)) x = 0; res &= x; }
    for (int r = 1; r < m->world; ++r) send_all(m->socks[r], &res, 4);
  } else {
    if (send_all(m->socks[0], &v, 4) || recv_all(m->socks[0], &res, 4)) res = 0;
  }
  *all_ok = res;
  if (!res && ok) return set_error(c, VCT_ERR_STATE, "vct_comm_init: a peer failed during set-up");
  return VCT_OK;
}

static void close_socks(Comm* m) {
  for (int s : m->socks) if (s >= 0) close(s);
  m->socks.clear();
}

// ------------------------------------------------------------------------------------------ set-up / tear-down
static CUmemAllocationProp alloc_prop(int device, bool exportable) {
  CUmemAllocationProp p = {};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = exportable ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
  return p;
}

static int map_rw(vct_context* c, CUdeviceptr va, size_t bytes, CUmemGenericAllocationHandle h, int device) {
  VCT_CU(c, g_drv.MemMap(va, bytes, 0, h, 0));
  CUmemAccessDesc d = {};
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE; d.location.id = device;
  d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  VCT_CU(c, g_drv.MemSetAccess(va, bytes, &d, 1));
  return VCT_OK;
}

static void comm_release(vct_context* c) {
  Comm* m = (Comm*)c->comm;
  if (!m) return;
  cudaSetDevice(c->device);
  sync_all_streams(c);
  if (g_drv.ok) {
    if (m->mc_va) { g_drv.MemUnmap(m->mc_va, m->seg); g_drv.MemAddressFree(m->mc_va, m->seg); }
    if (m->va) {
      for (int r = 0; r < m->world; ++r) if (r < (int)m->peers.size() && m->peers[r]) g_drv.MemUnmap(m->va + (size_t)r * m->seg, m->seg);
      g_drv.MemAddressFree(m->va, m->seg * m->world);
    }
    if (m->mc) {
      if (m->local && m->multicast) g_drv.MulticastUnbind(m->mc, m->cu_dev, 0, m->seg);
      if (!m->in_process) g_drv.MemRelease(m->mc);
    }
    if (m->group && --m->group->refs == 0) {
      if (m->group->mc) g_drv.MemRelease(m->group->mc);
      delete m->group;
    }
    m->group = nullptr;
    for (int r = 0; r < (int)m->peers.size(); ++r)
      if (m->peers[r] && (r != m->rank) && !m->in_process) g_drv.MemRelease(m->peers[r]);
    if (m->local) g_drv.MemRelease(m->local);
  }
  for (int k = 0; k < 3; ++k) { if (m->ev_band_done[k]) cudaEventDestroy(m->ev_band_done[k]); if (m->ev_copied[k]) cudaEventDestroy(m->ev_copied[k]); }
  close_socks(m);
  delete m;
  c->comm = nullptr;
  c->shared_local = nullptr; c->shared_mc = nullptr; c->shared_peers = nullptr; c->shared_seg = 0;
}

static int segment_size(vct_context* c, Comm* m, bool want_mc, size_t* seg) {
  m->inbox_bytes = align_up(inbox_bytes_for(c), 4096);
  m->frame_bytes = align_up((size_t)c->P.W * c->P.H * 4, 4096);
  m->off_inbox = COMM_PADS;
  m->off_frames = m->off_inbox + m->inbox_bytes;
  m->off_depth = m->off_frames + 3 * m->frame_bytes;
  m->depth_bytes = c->shard_shadow ? align_up((size_t)c->P.S * c->P.S * 4, 4096) : 0;     // sharded shadow map: one D24 image per rank
  size_t bytes = m->off_depth + m->depth_bytes;
  CUmemAllocationProp p = alloc_prop(c->device, !m->in_process);
  size_t gran = 0;
  VCT_CU(c, g_drv.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (want_mc) {
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)m->world; mp.size = align_up(bytes, gran);
    mp.handleTypes = m->in_process ? 0 : CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    VCT_CU(c, g_drv.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM));
    if (mg > gran) gran = mg;
  }
  *seg = align_up(bytes, gran);
  return VCT_OK;
}

static int finish_setup(vct_context* c, Comm* m) {
  // views for the exchange kernels (vct_voxelize.cu): own inbox, multicast inbox, base of the peer mappings
  c->shared_local = (unsigned long long*)(m->va + (size_t)m->rank * m->seg + m->off_inbox);
  c->shared_mc = m->multicast ? (unsigned long long*)(m->mc_va + m->off_inbox) : nullptr;
  c->shared_peers = (unsigned char*)m->va + m->off_inbox;      // rank r's inbox = shared_peers + r * shared_seg
  c->shared_seg = m->seg;
  c->exchange_parity = 0;
  {
    const size_t n = (size_t)c->P.V * c->P.V * c->P.V;
    size_t cap = c->exchange_cap_user ? c->exchange_cap_user : 32 * (size_t)c->P.V * c->P.V;
    c->exchange_cap = cap < n ? cap : n;
  }
  if (!c->d_push_count) VCT_CUDA(c, cudaMalloc(&c->d_push_count, 128));     // no allocation inside a sharded frame
  m->V = c->P.V; m->W = c->P.W; m->H = c->P.H; m->exch_cap = c->exchange_cap; m->shared_exchange = c->shared_exchange;
  for (int k = 0; k < 3; ++k) {
    VCT_CUDA(c, cudaEventCreateWithFlags(&m->ev_band_done[k], cudaEventDisableTiming));
    VCT_CUDA(c, cudaEventCreateWithFlags(&m->ev_copied[k], cudaEventDisableTiming));
  }
  c->scene_epoch++;
  return VCT_OK;
}

static int comm_init_process(vct_context* c, int rank, int world, const char* session, int flags) {
  int rc = load_driver(c); if (rc) return rc;
  if (c->comm) comm_release(c);
  Comm* m = new Comm();
  c->comm = m;
  m->rank = rank; m->world = world;
  c->shared_world = world; c->shared_rank = rank;
  auto fail = [&](int code) { std::string keep = c->err; comm_release(c); c->err = keep; return code; };
  VCT_CU(c, g_drv.DeviceGet(&m->cu_dev, c->device));
  int mc_supported = 0;
  g_drv.DeviceGetAttribute(&mc_supported, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, m->cu_dev);
  bool want_mc = mc_supported && !(flags & VCT_COMM_NO_MULTICAST) && world > 1;
  if (world > 1) { rc = bootstrap_connect(c, m, session ? session : "default", 120.0); if (rc) return fail(rc); }
  // every rank must take the same decision: multicast only if every rank can
  int all = 1;
  rc = host_barrier(c, m, want_mc ? 1 : 0, &all);
  want_mc = want_mc && all; c->err.clear();
  rc = segment_size(c, m, want_mc, &m->seg); if (rc) return fail(rc);
  // sizes must agree (they derive from the uniforms): rank 0 announces, the others compare
  if (world > 1) {
    unsigned long long seg0 = m->seg;
    if (m->rank == 0) { for (int r = 1; r < world; ++r) send_all(m->socks[r], &seg0, 8); }
    else if (recv_all(m->socks[0], &seg0, 8)) seg0 = 0;
    int same = seg0 == m->seg;
    rc = host_barrier(c, m, same, &all);
    if (!all) { set_error(c, VCT_ERR_STATE, "vct_comm_init: ranks disagree on VoxelDimensions / screen size / MaxExchangeVoxels"); return fail(VCT_ERR_STATE); }
  }
  CUmemAllocationProp prop = alloc_prop(c->device, true);
  if ((rc = check_cu(c, g_drv.MemCreate(&m->local, m->seg, &prop, 0), "cuMemCreate"))) return fail(rc);
  m->peers.assign(world, 0);
  m->peers[rank] = m->local;
  int ok = 1;
  if (world > 1) {
    int my_fd = -1;
    if (check_cu(c, g_drv.MemExportToShareableHandle(&my_fd, m->local, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle")) ok = 0;
    std::vector<int> fds(world + 1, -1);      // [0..world) = segments, [world] = multicast object
    if (m->rank == 0) {
      fds[0] = my_fd;
      for (int r = 1; r < world && ok; ++r) if (recv_fds(m->socks[r], &fds[r], 1)) ok = 0;
      if (ok && want_mc) {
        CUmulticastObjectProp mp = {};
        mp.numDevices = (unsigned)world; mp.size = m->seg; mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        if (check_cu(c, g_drv.MulticastCreate(&m->mc, &mp), "cuMulticastCreate") ||
            check_cu(c, g_drv.MemExportToShareableHandle(&fds[world], m->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export multicast")) ok = 0;
      }
      const int n_send = want_mc ? world + 1 : world;
      for (int r = 1; r < world; ++r) {
        int okr = ok;
        send_all(m->socks[r], &okr, 4);
        if (ok && send_fds(m->socks[r], fds.data(), n_send)) ok = 0;
      }
    } else {
      if (!ok || send_fds(m->socks[0], &my_fd, 1)) ok = 0;
      int ok0 = 0;
      if (recv_all(m->socks[0], &ok0, 4) || !ok0) ok = 0;
      if (ok && recv_fds(m->socks[0], fds.data(), want_mc ? world + 1 : world)) ok = 0;
    }
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      if (check_cu(c, g_drv.MemImportFromShareableHandle(&m->peers[r], (void*)(intptr_t)fds[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle")) ok = 0;
    }
    if (ok && want_mc && rank != 0 &&
        check_cu(c, g_drv.MemImportFromShareableHandle(&m->mc, (void*)(intptr_t)fds[world], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import multicast")) ok = 0;
    for (int fd : fds) if (fd >= 0 && fd != my_fd) close(fd);
    if (my_fd >= 0) close(my_fd);
    rc = host_barrier(c, m, ok, &all); if (!all) return fail(VCT_ERR_STATE);
    if (want_mc) {
      ok = check_cu(c, g_drv.MulticastAddDevice(m->mc, m->cu_dev), "cuMulticastAddDevice") == 0;
      rc = host_barrier(c, m, ok, &all); if (!all) return fail(VCT_ERR_STATE);          // every device added before any bind
      ok = check_cu(c, g_drv.MulticastBindMem(m->mc, 0, m->local, 0, m->seg, 0), "cuMulticastBindMem") == 0;
      rc = host_barrier(c, m, ok, &all); if (!all) return fail(VCT_ERR_STATE);
    }
  }
  // map: one VA range of world segments (segment r = rank r's memory) + the multicast view
  if ((rc = check_cu(c, g_drv.MemAddressReserve(&m->va, m->seg * world, 0, 0, 0), "cuMemAddressReserve"))) return fail(rc);
  for (int r = 0; r < world; ++r)
    if ((rc = map_rw(c, m->va + (size_t)r * m->seg, m->seg, m->peers[r], c->device))) return fail(rc);
  if (want_mc) {
    if ((rc = check_cu(c, g_drv.MemAddressReserve(&m->mc_va, m->seg, 0, 0, 0), "cuMemAddressReserve (multicast)"))) return fail(rc);
    if ((rc = map_rw(c, m->mc_va, m->seg, m->mc, c->device))) return fail(rc);
    m->multicast = true;
  }
  if ((rc = check_cuda(c, cudaMemsetAsync((void*)(m->va + (size_t)rank * m->seg), 0, m->seg, c->stream), "zero segment"))) return fail(rc);
  if ((rc = check_cuda(c, cudaStreamSynchronize(c->stream), "sync"))) return fail(rc);
  rc = host_barrier(c, m, 1, &all); if (!all) return fail(VCT_ERR_STATE);               // nobody signals into a pad that is zeroed later
  if ((rc = finish_setup(c, m))) return fail(rc);
  return VCT_OK;
}

// One process, several devices (vct_create_multi): the same segments / multicast object / mappings, created directly
// from the allocation handles -- nothing to export, no sockets.
static int comm_init_in_process(vct_context** cs, int n, int flags) {
  vct_context* c0 = cs[0];
  int rc = load_driver(c0); if (rc) return rc;
  std::vector<Comm*> ms(n, nullptr);
  int mc_all = 1;
  for (int r = 0; r < n; ++r) { cudaSetDevice(cs[r]->device); if (cs[r]->comm) comm_release(cs[r]); }
  InProcGroup* group = new InProcGroup();
  group->refs = n;
  for (int r = 0; r < n; ++r) {
    vct_context* c = cs[r];
    cudaSetDevice(c->device);
    Comm* m = new Comm();
    m->group = group;
    c->comm = m; ms[r] = m;
    m->rank = r; m->world = n; m->in_process = true;
    c->shared_world = n; c->shared_rank = r;
    if ((rc = check_cu(c, g_drv.DeviceGet(&m->cu_dev, c->device), "cuDeviceGet"))) return rc;
    int sup = 0;
    g_drv.DeviceGetAttribute(&sup, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, m->cu_dev);
    mc_all &= sup;
    for (int q = 0; q < r; ++q) if (cs[q]->device == c->device) mc_all = 0;      // one device twice: no multicast team
  }
  const bool want_mc = mc_all && n > 1 && !(flags & VCT_COMM_NO_MULTICAST);
  size_t seg = 0;
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(cs[r]->device);
    size_t s = 0;
    if ((rc = segment_size(cs[r], ms[r], want_mc, &s))) return rc;
    if (r && s != seg) return set_error(c0, VCT_ERR_STATE, "vct_comm_init_multi: the handles disagree on VoxelDimensions / screen size");
    seg = s;
  }
  CUmemGenericAllocationHandle mc = 0;
  if (want_mc) {
    CUmulticastObjectProp mp = {};
    mp.numDevices = (unsigned)n; mp.size = seg; mp.handleTypes = 0;
    if ((rc = check_cu(c0, g_drv.MulticastCreate(&mc, &mp), "cuMulticastCreate"))) return rc;
    group->mc = mc;
    for (int r = 0; r < n; ++r)
      if ((rc = check_cu(c0, g_drv.MulticastAddDevice(mc, ms[r]->cu_dev), "cuMulticastAddDevice"))) return rc;
  }
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(cs[r]->device);
    ms[r]->seg = seg; ms[r]->mc = mc;
    CUmemAllocationProp prop = alloc_prop(cs[r]->device, false);
    if ((rc = check_cu(cs[r], g_drv.MemCreate(&ms[r]->local, seg, &prop, 0), "cuMemCreate"))) return rc;
    if (want_mc && (rc = check_cu(cs[r], g_drv.MulticastBindMem(mc, 0, ms[r]->local, 0, seg, 0), "cuMulticastBindMem"))) return rc;
  }
  for (int r = 0; r < n; ++r) {
    vct_context* c = cs[r]; Comm* m = ms[r];
    cudaSetDevice(c->device);
    m->peers.assign(n, 0);
    for (int q = 0; q < n; ++q) m->peers[q] = ms[q]->local;
    if ((rc = check_cu(c, g_drv.MemAddressReserve(&m->va, seg * n, 0, 0, 0), "cuMemAddressReserve"))) return rc;
    for (int q = 0; q < n; ++q)
      if ((rc = map_rw(c, m->va + (size_t)q * seg, seg, m->peers[q], c->device))) return rc;
    if (want_mc) {
      if ((rc = check_cu(c, g_drv.MemAddressReserve(&m->mc_va, seg, 0, 0, 0), "cuMemAddressReserve (multicast)"))) return rc;
      if ((rc = map_rw(c, m->mc_va, seg, mc, c->device))) return rc;
      m->multicast = true;
    }
    if ((rc = check_cuda(c, cudaMemsetAsync((void*)(m->va + (size_t)r * seg), 0, seg, c->stream), "zero segment"))) return rc;
    if ((rc = check_cuda(c, cudaStreamSynchronize(c->stream), "sync"))) return rc;
  }
  for (int r = 0; r < n; ++r) {
    cudaSetDevice(cs[r]->device);
    if ((rc = finish_setup(cs[r], ms[r]))) return rc;
  }
  return VCT_OK;
}

// enqueue the device barrier of `channel` on `stream`
int comm_barrier(vct_context* c, int channel, cudaStream_t stream) {
  Comm* m = (Comm*)c->comm;
  if (!m || m->world == 1) return VCT_OK;
  const uint32_t e = ++m->epoch[channel];
  comm_barrier_kernel<<<1, 32, 0, stream>>>((uint32_t*)m->va, m->seg / 4, m->rank, m->world, channel, e, 10ull * 1000 * 1000 * 1000);
  c->launches += 1;
  return check_cuda(c, cudaGetLastError(), "comm_barrier");
}

int comm_check(vct_context* c) {
  Comm* m = (Comm*)c->comm;
  if (!m || m->world == 1) return VCT_OK;
  uint32_t flag = 0;
  VCT_CUDA(c, cudaMemcpy(&flag, (const void*)(m->va + (size_t)m->rank * m->seg + 2048), 4, cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset((void*)(m->va + (size_t)m->rank * m->seg + 2048), 0, 4);
    return set_error(c, VCT_ERR_STATE, "multi-GPU barrier timed out waiting for rank " + std::to_string((int)flag - 1));
  }
  return VCT_OK;
}

static int comm_settings_match(vct_context* c, Comm* m) {
  if (m->V != c->P.V || m->W != c->P.W || m->H != c->P.H || m->shared_exchange != c->shared_exchange)
    return set_error(c, VCT_ERR_STATE, "VoxelDimensions / screen size changed after vct_comm_init: call vct_comm_init again");
  return VCT_OK;
}

// the symmetric D24 image of the sharded shadow map: local view, multicast view (null without a multicast object), and
// the base + stride of the peer views
bool comm_depth_views(vct_context* c, uint32_t** local, uint32_t** mc, uint32_t** peers, size_t* seg_words, int* world, int* rank) {
  Comm* m = (Comm*)c->comm;
  if (!m || !m->depth_bytes || m->depth_bytes < (size_t)c->P.S * c->P.S * 4) return false;
  *local = (uint32_t*)(m->va + (size_t)m->rank * m->seg + m->off_depth);
  *mc = m->multicast ? (uint32_t*)(m->mc_va + m->off_depth) : nullptr;
  *peers = (uint32_t*)(m->va + m->off_depth);
  *seg_words = m->seg / 4; *world = m->world; *rank = m->rank;
  return true;
}

uchar4* comm_frame_slot(vct_context* c, int rank, int slot) {
  Comm* m = (Comm*)c->comm;
  return (uchar4*)(m->va + (size_t)rank * m->seg + m->off_frames + (size_t)slot * m->frame_bytes);
}

static void default_shares(vct_context* c, int rank, int world, int flags) {
  // triangles dealt in blocks of 128 round-robin; rows dealt in strips of 8 (cone_trace's block height) round-robin,
  // or -- VCT_COMM_ROW_BANDS -- as equal contiguous bands (multiples of 8 rows)
  c->tri_interleave = world; c->tri_phase = rank;
  c->P.row_begin = 0; c->P.row_end = 0; c->P.row_il = 0; c->P.row_ph = 0;
  if (world > 1 && (flags & VCT_COMM_ROW_BANDS)) {
    const int blocks = (c->P.H + 7) / 8, per = ((blocks + world - 1) / world) * 8;
    c->P.row_begin = rank * per < c->P.H ? rank * per : c->P.H;
    c->P.row_end = (rank + 1) * per < c->P.H ? (rank + 1) * per : c->P.H;
  } else if (world > 1) {
    c->P.row_il = world; c->P.row_ph = rank;
  }
  c->scene_epoch++;
}

}  // namespace vct

using namespace vct;

#define NEED(c) do { if (!(c)) return VCT_ERR_INVALID; cudaSetDevice((c)->device); } while (0)

extern "C" {

int vct_comm_init(vct_handle c, int rank, int world, const char* session, int flags) {
  NEED(c);
  if (world < 1 || world > 16 || rank < 0 || rank >= world) return set_error(c, VCT_ERR_INVALID, "vct_comm_init: bad rank / world");
  if (c->shared_frame_open) return set_error(c, VCT_ERR_STATE, "vct_comm_init: a shared frame is open");
  int rc = sync_all_streams(c); if (rc) return rc;
  rc = comm_init_process(c, rank, world, session, flags); if (rc) return rc;
  if (!(flags & VCT_COMM_KEEP_SHARES)) default_shares(c, rank, world, flags);
  return VCT_OK;
}

int vct_comm_destroy(vct_handle c) {
  NEED(c);
  comm_release(c);
  c->shared_world = 1; c->shared_rank = 0;
  return VCT_OK;
}

int vct_comm_info(vct_handle c, int* rank, int* world, int* multicast, size_t* segment_bytes) {
  NEED(c);
  Comm* m = (Comm*)c->comm;
  if (!m) return set_error(c, VCT_ERR_STATE, "vct_comm_info: vct_comm_init has not been called");
  if (rank) *rank = m->rank;
  if (world) *world = m->world;
  if (multicast) *multicast = m->multicast ? 1 : 0;
  if (segment_bytes) *segment_bytes = m->seg;
  return VCT_OK;
}

int vct_comm_barrier(vct_handle c) {
  NEED(c);
  if (!c->comm) return set_error(c, VCT_ERR_STATE, "vct_comm_barrier: vct_comm_init has not been called");
  return comm_barrier(c, 2, c->stream);
}

// One sharded frame, entirely inside the library (replaces main.cpp:81-92 for a multi-GPU loop):
//   voxel stream : vertex pass, sparse clear, cover + shade of this rank's triangle share, push (multimem.st),
//                  BARRIER 0, merge, resolve, mip                      (beside cone_trace of the previous frame)
//   visibility   : primary visibility of this rank's rows
//   main stream  : cone_trace of this rank's rows, written straight into RANK 0's frame ring over NVLink (the gather
//                  is cone_trace's own epilogue: no staging copy, no collective), BARRIER 1; rank 0 then queues the
//                  device->host copy of the assembled frame on its copy stream.
// Split in two phases so that one host thread can drive several devices (vct_frame_sharded_multi).
static int sharded_phase_a(vct_context* c) {
  Comm* m = (Comm*)c->comm;
  if (!m) return set_error(c, VCT_ERR_STATE, "vct_frame_sharded: call vct_comm_init first");
  int rc = comm_settings_match(c, m); if (rc) return rc;
  const int slot = (int)(m->frame_seq % 3);
  if (m->copy_pending[slot]) {              // the host copy that last used this ring slot (three frames ago)
    VCT_CUDA(c, cudaEventSynchronize(m->ev_copied[slot]));
    m->copy_pending[slot] = false;
  }
  rc = vct_frame_shared_begin(c, 0, c->nt); if (rc) return rc;
  return comm_barrier(c, 0, c->stream_vox);
}

static int sharded_phase_b(vct_context* c, uint8_t* host_rgba) {
  Comm* m = (Comm*)c->comm;
  const int slot = (int)(m->frame_seq % 3);
  uchar4* saved = c->d_frame;
  c->d_frame = comm_frame_slot(c, 0, slot);                 // every rank writes its rows into rank 0's slot
  int rc = vct_frame_shared_end(c, nullptr);
  c->d_frame = saved;
  if (rc) return rc;
  // rank 0 may not let the peers overwrite the NEXT slot before its previous host copy has drained
  const int next = (slot + 1) % 3;
  if (m->rank == 0 && m->copy_pending[next]) VCT_CUDA(c, cudaStreamWaitEvent(c->stream, m->ev_copied[next], 0));
  rc = comm_barrier(c, 1, c->stream); if (rc) return rc;
  VCT_CUDA(c, cudaEventRecord(m->ev_band_done[slot], c->stream));
  if (m->rank == 0 && host_rgba) {
    if (!c->copy_stream) VCT_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    VCT_CUDA(c, cudaStreamWaitEvent(c->copy_stream, m->ev_band_done[slot], 0));
    VCT_CUDA(c, cudaMemcpyAsync(host_rgba, comm_frame_slot(c, 0, slot), (size_t)c->P.W * c->P.H * 4, cudaMemcpyDeviceToHost, c->copy_stream));
    VCT_CUDA(c, cudaEventRecord(m->ev_copied[slot], c->copy_stream));
    m->copy_pending[slot] = true;
  }
  m->frame_seq++;
  return VCT_OK;
}

// Returns without waiting; vct_frame_sharded_wait blocks until the frames in flight are complete (rank 0: in host memory).
int vct_frame_sharded(vct_handle c, uint8_t* host_rgba) {
  NEED(c);
  int rc = sharded_phase_a(c); if (rc) return rc;
  return sharded_phase_b(c, host_rgba);
}

// ---- one process, several devices
int vct_create_multi(const int* devices, int n, vct_handle* out) {
  if (!devices || !out || n < 1 || n > 16) return VCT_ERR_INVALID;
  for (int r = 0; r < n; ++r) out[r] = nullptr;
  for (int r = 0; r < n; ++r) {
    int rc = vct_create(devices[r], &out[r]);
    if (rc) { for (int q = 0; q < r; ++q) { vct_destroy(out[q]); out[q] = nullptr; } return rc; }
  }
  return VCT_OK;
}

int vct_comm_init_multi(vct_handle* hs, int n, int flags) {
  if (!hs || n < 1 || n > 16) return VCT_ERR_INVALID;
  for (int r = 0; r < n; ++r) {
    if (!hs[r]) return VCT_ERR_INVALID;
    cudaSetDevice(hs[r]->device);
    if (hs[r]->shared_frame_open) return set_error(hs[r], VCT_ERR_STATE, "vct_comm_init_multi: a shared frame is open");
    int rc = sync_all_streams(hs[r]); if (rc) return rc;
  }
  int rc = comm_init_in_process(hs, n, flags);
  if (rc) {
    std::string msg;
    for (int r = 0; r < n; ++r) if (!hs[r]->err.empty()) { msg = hs[r]->err; break; }
    for (int r = 0; r < n; ++r) { cudaSetDevice(hs[r]->device); comm_release(hs[r]); hs[r]->err = msg; }
    return rc;
  }
  if (!(flags & VCT_COMM_KEEP_SHARES)) for (int r = 0; r < n; ++r) default_shares(hs[r], r, n, flags);
  return VCT_OK;
}

int vct_frame_sharded_multi(vct_handle* hs, int n, uint8_t* host_rgba) {
  if (!hs || n < 1) return VCT_ERR_INVALID;
  // First frame: let every handle render one private frame so that all lazy allocations happen now.  cudaMalloc
  // synchronises its device; if two handles share a device, an allocation made while the other handle's barrier
  // kernel is already spinning would wait for a peer that this same host thread has not enqueued yet.
  for (int r = 0; r < n; ++r) {
    Comm* m = hs[r] ? (Comm*)hs[r]->comm : nullptr;
    if (!m) return hs[r] ? set_error(hs[r], VCT_ERR_STATE, "vct_frame_sharded_multi: call vct_comm_init_multi first") : VCT_ERR_INVALID;
    if (m->frame_seq == 0) {
      cudaSetDevice(hs[r]->device);
      int rc = vct_frame(hs[r], nullptr); if (rc) return rc;
      rc = sync_all_streams(hs[r]); if (rc) return rc;
    }
  }
  for (int r = 0; r < n; ++r) { cudaSetDevice(hs[r]->device); int rc = sharded_phase_a(hs[r]); if (rc) return rc; }
  // rank 0 last: with a pageable host buffer its cudaMemcpyAsync blocks the host until the frame is assembled, which
  // needs every other rank's second barrier to be enqueued already
  for (int r = n - 1; r >= 0; --r) { cudaSetDevice(hs[r]->device); int rc = sharded_phase_b(hs[r], r == 0 ? host_rgba : nullptr); if (rc) return rc; }
  return VCT_OK;
}

// blocks until every sharded frame issued so far is complete on this rank (rank 0: including its host copies)
int vct_frame_sharded_wait(vct_handle c) {
  NEED(c);
  Comm* m = (Comm*)c->comm;
  if (!m) return set_error(c, VCT_ERR_STATE, "vct_frame_sharded_wait: call vct_comm_init first");
  VCT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; ++k)
    if (m->copy_pending[k]) { VCT_CUDA(c, cudaEventSynchronize(m->ev_copied[k])); m->copy_pending[k] = false; }
  int rc = comm_check(c); if (rc) return rc;
  return check_overflow(c);
}

// the assembled frame of the most recent vct_frame_sharded (rank 0; other ranks see their own rows only)
int vct_comm_frame_buffer(vct_handle c, void** device_ptr, size_t* n_bytes) {
  NEED(c);
  Comm* m = (Comm*)c->comm;
  if (!m || !m->frame_seq) return set_error(c, VCT_ERR_STATE, "vct_comm_frame_buffer: no sharded frame rendered");
  if (device_ptr) *device_ptr = comm_frame_slot(c, m->rank == 0 ? 0 : 0, (int)((m->frame_seq - 1) % 3));
  if (n_bytes) *n_bytes = (size_t)c->P.W * c->P.H * 4;
  return VCT_OK;
}

}  // extern "C"

namespace vct {
void comm_release_for_destroy(vct_context* c) { comm_release(c); }
}
