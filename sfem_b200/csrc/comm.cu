// Distributed half of the C ABI (include/b200fem.h, "Communicator and halo exchange"): one process per GPU.
//
//   b200fem_comm_*  : a communicator over NCCL (ncclCommInitRank from a 128-byte unique id the launcher hands round),
//                     device-scalar allreduce for the CG / norm reductions (comm->sum of
//                     algebra/parallel/sfem_parallel_cg.hpp:169-172), barrier.
//   b200fem_halo_*  : smesh::Exchange::gather for the [owned | ghosts by owner] node layout of
//                     frontend/sfem_ParallelMatrixFreeOperator.cpp:78-105. Two transports:
//       PUSH (default): every rank owns a receive buffer (two parities) + epoch flags in ONE cudaMalloc block whose
//                     cudaIpc handle the peers map. gather(x) is two launches: (1) halo_push_kernel -- per destination peer
//                     the OWNER packs x[send_idx] straight into that peer's receive buffer (pack and NVLink transfer are
//                     one kernel of coalesced remote stores), fences, and the last block releases the peer's `arrive`
//                     epoch; (2) halo_wait_copy_kernel -- spins on its own `arrive` epochs, copies the received block
//                     into the contiguous ghost range of x and releases the `credit` epoch that allows the owner to
//                     reuse that parity two gathers later. Any device vector with ghost room works (nothing about x
//                     has to be symmetric), no collective happens after creation, and every spin has a time-out that
//                     raises the communicator's error flag instead of hanging the GPU.
//       NCCL        : pack kernel per peer + one group of ncclSend / ncclRecv straight into the ghost ranges.
//
// NCCL is resolved at run time (dlopen of the copy already in the process -- e.g. the one torch loaded -- else the
// system libnccl.so.2), so the library still loads on a box without NCCL and single-rank runs never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "common.cuh"

namespace b200fem {

// ---------------------------------------------------------------------------------------------------------------
// NCCL entry points, resolved lazily
// ---------------------------------------------------------------------------------------------------------------
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    std::string why;
};

static NcclApi *nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return &api;
    tried = true;
    void *h = nullptr;
    const char *env = getenv("B200FEM_NCCL_LIB");
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy this process already uses (torch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.why = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "?");
        return &api;
    }
#define B2_SYM(field, name)                                                          \
    do {                                                                             \
        *(void **)(&api.field) = dlsym(h, name);                                     \
        if (!api.field) {                                                            \
            api.why = std::string("libnccl is missing ") + name;                     \
            return &api;                                                             \
        }                                                                            \
    } while (0)
    B2_SYM(GetUniqueId, "ncclGetUniqueId");
    B2_SYM(CommInitRank, "ncclCommInitRank");
    B2_SYM(CommDestroy, "ncclCommDestroy");
    B2_SYM(AllReduce, "ncclAllReduce");
    B2_SYM(AllGather, "ncclAllGather");
    B2_SYM(Send, "ncclSend");
    B2_SYM(Recv, "ncclRecv");
    B2_SYM(GroupStart, "ncclGroupStart");
    B2_SYM(GroupEnd, "ncclGroupEnd");
    B2_SYM(GetErrorString, "ncclGetErrorString");
#undef B2_SYM
    api.ok = true;
    return &api;
}

#define B2_NCCL(call)                                                                                         \
    do {                                                                                                      \
        ncclResult_t r_ = (call);                                                                             \
        if (r_ != ncclSuccess) {                                                                              \
            set_error(std::string("NCCL error '") + nccl_api()->GetErrorString(r_) + "' in " #call);         \
            return B200FEM_FAILURE;                                                                           \
        }                                                                                                     \
    } while (0)

}  // namespace b200fem

using namespace b200fem;

struct b200fem_comm_s {
    ncclComm_t nccl = nullptr;
    int rank = 0, world = 1, device = 0;
    double *dev_scratch = nullptr;   // [kScratch] staging for host-valued reductions
    double *host_scratch = nullptr;  // pinned
    int *err_host = nullptr;         // pinned + mapped: raised by a spin that timed out
    int *err_dev = nullptr;
    unsigned long long timeout_ns = 60ull * 1000000000ull;
    static constexpr int kScratch = 64;
};

namespace b200fem {

// ---------------------------------------------------------------------------------------------------------------
// PUSH transport kernels
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
constexpr int kHaloThreads = 256;
constexpr size_t kFlagBytes = 4096;  // arrive[2][world] + credit[2][world] uint64, world <= 64

struct PushPeer {
    const int32_t *idx;                 // local node ids the peer needs (device)
    ptrdiff_t n;                        // nodes
    void *remote_data[2];               // where my block starts in the peer's receive buffer, per parity
    unsigned long long *remote_arrive[2];     // peer's arrive[parity][me]
    const unsigned long long *credit[2];      // my credit[parity][peer]
    unsigned int *ticket;
};
struct RecvPeer {
    const void *data[2];                // my receive buffer block for this peer, per parity
    ptrdiff_t n_values;                 // nodes * block_size
    ptrdiff_t dst_offset;               // first value of the ghost range in x
    const unsigned long long *arrive[2];      // my arrive[parity][peer]
    unsigned long long *remote_credit[2];     // peer's credit[parity][me]
    unsigned int *ticket;
    const int32_t *idx;                 // indexed receive: local node id of every received node (NULL: the contiguous range)
};
struct PushArgs {
    PushPeer p[kMaxPeers];
};
struct RecvArgs {
    RecvPeer p[kMaxPeers];
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// thread 0 of the block waits until *flag >= want (acquire, system scope); gives up after timeout_ns
__device__ __forceinline__ void block_wait_epoch(const unsigned long long *flag, unsigned long long want,
                                                 unsigned long long timeout_ns, int *err) {
    if (threadIdx.x == 0 && want > 0) {
        if (ld_acquire_sys(flag) < want) {
            const unsigned long long t0 = globaltimer_ns();
            unsigned int spins = 0;
            while (ld_acquire_sys(flag) < want) {
                __nanosleep(64);
                if ((++spins & 1023u) == 0) {
                    // a peer that never arrives must not hang the GPU: raise the communicator's error flag (the host
                    // reads it at the next reduction) and let every later wait fall through at once
                    if (*((volatile int *)err) != 0) break;
                    if (globaltimer_ns() - t0 > timeout_ns) {
                        *((volatile int *)err) = 1;
                        __threadfence_system();
                        break;
                    }
                }
            }
        }
    }
    __syncthreads();
}

// last block of (grid.x) to get here releases `flag = epoch` at system scope
__device__ __forceinline__ void block_release_epoch(unsigned int *ticket, unsigned long long *flag, unsigned long long epoch) {
    __threadfence_system();  // every thread: my stores are visible system-wide before the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence_system();
            st_release_sys(flag, epoch);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kHaloThreads)
    halo_push_kernel(const PushArgs a, const int bs, const T *__restrict__ x, const unsigned long long epoch,
                     const unsigned long long timeout_ns, int *err) {
    const PushPeer &p = a.p[blockIdx.y];
    const int par = (int)(epoch & 1ull);
    // the peer has drained this parity's buffer of gather epoch - 2
    block_wait_epoch(p.credit[par], epoch >= 2 ? epoch - 2 : 0, timeout_ns, err);
    T *dst = reinterpret_cast<T *>(p.remote_data[par]);
    const ptrdiff_t total = p.n * bs;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t node = i / bs;
        const int c = (int)(i - node * bs);
        dst[i] = x[(ptrdiff_t)p.idx[node] * bs + c];
    }
    block_release_epoch(p.ticket, p.remote_arrive[par], epoch);
}

template <typename T>
__global__ void __launch_bounds__(kHaloThreads)
    halo_wait_copy_kernel(const RecvArgs a, const int bs, T *__restrict__ x, const unsigned long long epoch,
                          const unsigned long long timeout_ns, int *err, unsigned long long *ready, unsigned int *ready_ticket) {
    const RecvPeer &p = a.p[blockIdx.y];
    const int par = (int)(epoch & 1ull);
    block_wait_epoch(p.arrive[par], epoch, timeout_ns, err);
    const T *src = reinterpret_cast<const T *>(p.data[par]);
    T *dst = x + p.dst_offset;
    if (p.idx) {  // indexed receive: the values of the requested nodes only, scattered to their places
        for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_values; i += (ptrdiff_t)gridDim.x * blockDim.x) {
            const ptrdiff_t node = i / bs;
            x[(ptrdiff_t)p.idx[node] * bs + (int)(i - node * bs)] = __ldcg(&src[i]);
        }
    } else
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n_values; i += (ptrdiff_t)gridDim.x * blockDim.x)
        dst[i] = __ldcg(&src[i]);  // L2 only: the lines were written by the peer, never trust this SM's L1
    block_release_epoch(p.ticket, p.remote_credit[par], epoch);
    // deferred form: the last block of the whole grid tells the apply kernel that spins on `ready` (same GPU) that the ghost
    // block of x is complete (every block's stores were fenced in block_release_epoch)
    if (ready && threadIdx.x == 0) {
        const unsigned int t = atomicAdd(ready_ticket, 1u);
        if (t == gridDim.x * gridDim.y - 1) {
            *ready_ticket = 0u;
            __threadfence();
            asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(ready), "l"(epoch) : "memory");
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kHaloThreads)
    halo_pack_kernel(ptrdiff_t n, const int32_t *__restrict__ idx, int bs, const T *__restrict__ src, T *__restrict__ dst) {
    const ptrdiff_t total = n * bs;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t node = i / bs;
        dst[i] = src[(ptrdiff_t)idx[node] * bs + (int)(i - node * bs)];
    }
}

template <typename T>
__global__ void __launch_bounds__(kHaloThreads)
    halo_unpack_kernel(ptrdiff_t n, const int32_t *__restrict__ idx, int bs, const T *__restrict__ src, T *__restrict__ dst) {
    const ptrdiff_t total = n * bs;
    for (ptrdiff_t i = (ptrdiff_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (ptrdiff_t)gridDim.x * blockDim.x) {
        const ptrdiff_t node = i / bs;
        dst[(ptrdiff_t)idx[node] * bs + (int)(i - node * bs)] = src[i];
    }
}

static int comm_allgather_bytes(b200fem_comm_s *c, const void *mine, size_t bytes, std::vector<unsigned char> *all) {
    all->assign(bytes * (size_t)c->world, 0);
    if (c->world == 1) {
        memcpy(all->data(), mine, bytes);
        return B200FEM_SUCCESS;
    }
    unsigned char *d_in = nullptr, *d_out = nullptr;
    B2_CUDA(cudaMalloc((void **)&d_in, bytes));
    B2_CUDA(cudaMalloc((void **)&d_out, bytes * (size_t)c->world));
    B2_CUDA(cudaMemcpy(d_in, mine, bytes, cudaMemcpyHostToDevice));
    B2_NCCL(nccl_api()->AllGather(d_in, d_out, bytes, ncclUint8, c->nccl, (cudaStream_t)0));
    B2_CUDA(cudaStreamSynchronize((cudaStream_t)0));
    B2_CUDA(cudaMemcpy(all->data(), d_out, bytes * (size_t)c->world, cudaMemcpyDeviceToHost));
    cudaFree(d_in);
    cudaFree(d_out);
    return B200FEM_SUCCESS;
}

}  // namespace b200fem

struct b200fem_halo_s {
    b200fem_comm_s *comm = nullptr;
    int real_type = B200FEM_F64, bs = 1, transport = B200FEM_HALO_NCCL;
    ptrdiff_t n_owned = 0, n_local = 0;
    std::vector<int> send_peers, recv_peers;
    std::vector<ptrdiff_t> send_count, recv_begin, recv_count;
    std::vector<DeviceArray<int32_t>> send_idx;
    std::vector<DeviceArray<int32_t>> recv_idx;  // indexed receive (b200fem_halo_set_recv_index), empty otherwise
    // NCCL transport
    std::vector<void *> send_buf, recv_stage;  // recv_stage: indexed receive only
    // PUSH transport
    unsigned char *block = nullptr;  // [flags | parity 0 | parity 1]
    size_t parity_stride = 0;
    std::vector<void *> peer_block;  // mapped blocks of the peers (by rank), nullptr when unused
    unsigned int *tickets = nullptr;
    PushArgs push{};
    RecvArgs recv{};
    unsigned long long epoch = 0;
    long n_gathers = 0;
    // deferred receiver half (b200fem_halo_gather_end_deferred): its own stream, a ready flag + ticket in device memory
    cudaStream_t side = nullptr;
    cudaEvent_t side_event = nullptr;
    unsigned long long *ready = nullptr;  // [0] = epoch of the last complete ghost block, [1] = ticket
};

namespace b200fem {

static void halo_release(b200fem_halo_s *h) {
    if (!h) return;
    for (void *p : h->send_buf)
        if (p) cudaFree(p);
    for (void *p : h->recv_stage)
        if (p) cudaFree(p);
    for (size_t q = 0; q < h->peer_block.size(); ++q)
        if (h->peer_block[q]) cudaIpcCloseMemHandle(h->peer_block[q]);
    if (h->tickets) cudaFree(h->tickets);
    if (h->block) cudaFree(h->block);
    if (h->side) {
        cudaStreamSynchronize(h->side);
        cudaStreamDestroy(h->side);
    }
    if (h->side_event) cudaEventDestroy(h->side_event);
    if (h->ready) cudaFree(h->ready);
    delete h;
}

// Receive block + flags, handle exchange, descriptor tables. Returns FAILURE (with the error set) when peer mapping is
// not possible on this box; the caller then falls back to NCCL.
static int halo_setup_push(b200fem_halo_s *h) {
    b200fem_comm_s *c = h->comm;
    const int W = c->world, me = c->rank;
    if ((int)h->send_peers.size() > kMaxPeers || (int)h->recv_peers.size() > kMaxPeers) B2_FAIL("halo: too many peers for the push transport");
    if ((size_t)W * 4 * sizeof(unsigned long long) > kFlagBytes) B2_FAIL("halo: world too large for the push transport");
    const size_t vsize = (size_t)h->real_type * (size_t)h->bs;
    const size_t ghost_bytes = (size_t)(h->n_local - h->n_owned) * vsize;
    h->parity_stride = (ghost_bytes + 255) / 256 * 256;
    const size_t block_bytes = kFlagBytes + 2 * h->parity_stride + 256;
    B2_CUDA(cudaMalloc((void **)&h->block, block_bytes));
    B2_CUDA(cudaMemset(h->block, 0, block_bytes));
    B2_CUDA(cudaMalloc((void **)&h->tickets, sizeof(unsigned int) * 2 * kMaxPeers));
    B2_CUDA(cudaMemset(h->tickets, 0, sizeof(unsigned int) * 2 * kMaxPeers));

    // what every rank publishes: its block's IPC handle, its parity stride, and per source rank the offset / count
    // (in nodes, relative to the ghost range) of the block it expects from that rank
    struct Meta {
        cudaIpcMemHandle_t handle;
        long long parity_stride;
        long long ok;
    };
    const size_t rec = sizeof(Meta) + sizeof(long long) * 2 * (size_t)W;
    std::vector<unsigned char> mine(rec, 0), all;
    Meta *m = reinterpret_cast<Meta *>(mine.data());
    long long *tab = reinterpret_cast<long long *>(mine.data() + sizeof(Meta));
    for (int q = 0; q < W; ++q) tab[2 * q] = tab[2 * q + 1] = -1;
    m->parity_stride = (long long)h->parity_stride;
    m->ok = (cudaIpcGetMemHandle(&m->handle, h->block) == cudaSuccess) ? 1 : 0;
    if (!m->ok) cudaGetLastError();
    for (size_t j = 0; j < h->recv_peers.size(); ++j) {
        tab[2 * h->recv_peers[j]] = (long long)(h->recv_begin[j] - h->n_owned);
        tab[2 * h->recv_peers[j] + 1] = (long long)h->recv_count[j];
    }
    if (comm_allgather_bytes(c, mine.data(), rec, &all)) return B200FEM_FAILURE;
    auto meta_of = [&](int q) { return reinterpret_cast<const Meta *>(all.data() + rec * (size_t)q); };
    auto tab_of = [&](int q) { return reinterpret_cast<const long long *>(all.data() + rec * (size_t)q + sizeof(Meta)); };

    // every rank must take the same decision: map peers only if everyone could export
    bool all_ok = true;
    for (int q = 0; q < W; ++q) all_ok = all_ok && meta_of(q)->ok;
    int mapped_ok = all_ok ? 1 : 0;
    h->peer_block.assign((size_t)W, nullptr);
    if (all_ok) {
        std::vector<int> need;
        for (int q : h->send_peers) need.push_back(q);
        for (int q : h->recv_peers) need.push_back(q);
        std::sort(need.begin(), need.end());
        need.erase(std::unique(need.begin(), need.end()), need.end());
        for (int q : need) {
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, meta_of(q)->handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                mapped_ok = 0;
                break;
            }
            h->peer_block[(size_t)q] = p;
        }
    }
    {  // agree on the outcome (a rank that failed to map forces NCCL everywhere)
        std::vector<unsigned char> votes;
        long long v = mapped_ok;
        if (comm_allgather_bytes(c, &v, sizeof(v), &votes)) return B200FEM_FAILURE;
        for (int q = 0; q < W; ++q) mapped_ok = mapped_ok && (reinterpret_cast<const long long *>(votes.data())[q] != 0);
    }
    if (!mapped_ok) B2_FAIL("halo: cudaIpc peer mapping unavailable");

    auto flags = [&](unsigned char *base, int which /*0 arrive, 1 credit*/, int par, int r) {
        return reinterpret_cast<unsigned long long *>(base) + ((size_t)(which * 2 + par) * (size_t)W + (size_t)r);
    };
    for (size_t j = 0; j < h->send_peers.size(); ++j) {
        const int q = h->send_peers[j];
        const long long off = tab_of(q)[2 * me], cnt = tab_of(q)[2 * me + 1];
        if (off < 0 || cnt != (long long)h->send_count[j]) B2_FAIL("halo: send list does not match the peer's receive range");
        unsigned char *pb = (unsigned char *)h->peer_block[(size_t)q];
        PushPeer &pp = h->push.p[j];
        pp.idx = h->send_idx[j].ptr;
        pp.n = h->send_count[j];
        for (int par = 0; par < 2; ++par) {
            pp.remote_data[par] = pb + kFlagBytes + (size_t)par * (size_t)meta_of(q)->parity_stride + (size_t)off * vsize;
            pp.remote_arrive[par] = flags(pb, 0, par, me);
            pp.credit[par] = flags(h->block, 1, par, q);
        }
        pp.ticket = h->tickets + j;
    }
    for (size_t j = 0; j < h->recv_peers.size(); ++j) {
        const int p = h->recv_peers[j];
        unsigned char *pb = (unsigned char *)h->peer_block[(size_t)p];
        RecvPeer &rp = h->recv.p[j];
        rp.n_values = h->recv_count[j] * h->bs;
        rp.dst_offset = h->recv_begin[j] * h->bs;
        for (int par = 0; par < 2; ++par) {
            rp.data[par] = h->block + kFlagBytes + (size_t)par * h->parity_stride + (size_t)(h->recv_begin[j] - h->n_owned) * vsize;
            rp.arrive[par] = flags(h->block, 0, par, p);
            rp.remote_credit[par] = flags(pb, 1, par, me);
        }
        rp.ticket = h->tickets + kMaxPeers + j;
    }
    return B200FEM_SUCCESS;
}

// phase: 0 = the whole gather, 1 = begin (the owner side: push / pack + send), 2 = end (the receiver side: wait + copy).
// The NCCL transport is one grouped call and runs entirely in `begin`.
template <typename T>
static int halo_gather_t(b200fem_halo_s *h, T *x, cudaStream_t s, int phase = 0) {
    b200fem_comm_s *c = h->comm;
    const bool deferred = phase == 3;  // the whole gather on the halo's own stream, announcing through h->ready (epoch set by _begin)
    if (deferred) phase = 0;
    if (phase != 2) h->n_gathers++;
    if (c->world == 1 || (h->send_peers.empty() && h->recv_peers.empty())) return B200FEM_SUCCESS;
    if (h->transport == B200FEM_HALO_PUSH) {
        const unsigned long long epoch = (phase == 2 || deferred) ? h->epoch : ++h->epoch;
        ptrdiff_t max_send = 0, max_recv = 0;
        for (auto n : h->send_count) max_send = std::max(max_send, n * h->bs);
        for (size_t j = 0; j < h->recv_count.size(); ++j) max_recv = std::max(max_recv, h->recv_count[j] * (ptrdiff_t)h->bs);
        const ptrdiff_t cap = (ptrdiff_t)sm_count() * 4;
        if (phase != 2 && !h->send_peers.empty()) {
            const int gx = (int)std::max<ptrdiff_t>(1, std::min<ptrdiff_t>((max_send + kHaloThreads * 4 - 1) / (kHaloThreads * 4),
                                                                          cap / (ptrdiff_t)h->send_peers.size()));
            halo_push_kernel<T><<<dim3(gx, (unsigned)h->send_peers.size()), kHaloThreads, 0, deferred ? h->side : s>>>(
                    h->push, h->bs, x, epoch, c->timeout_ns, c->err_dev);
            B2_CHECK_LAUNCH();
        }
        if (phase != 1 && !h->recv_peers.empty()) {
            const int gx = (int)std::max<ptrdiff_t>(1, std::min<ptrdiff_t>((max_recv + kHaloThreads * 4 - 1) / (kHaloThreads * 4),
                                                                          cap / (ptrdiff_t)h->recv_peers.size()));
            if (deferred) {
                halo_wait_copy_kernel<T><<<dim3(gx, (unsigned)h->recv_peers.size()), kHaloThreads, 0, h->side>>>(
                        h->recv, h->bs, x, epoch, c->timeout_ns, c->err_dev, h->ready, reinterpret_cast<unsigned int *>(h->ready + 1));
            } else {
                halo_wait_copy_kernel<T><<<dim3(gx, (unsigned)h->recv_peers.size()), kHaloThreads, 0, s>>>(h->recv, h->bs, x, epoch, c->timeout_ns,
                                                                                                          c->err_dev, nullptr, nullptr);
            }
            B2_CHECK_LAUNCH();
        }
        return B200FEM_SUCCESS;
    }
    // NCCL transport
    if (phase == 2) return B200FEM_SUCCESS;
    const ncclDataType_t dt = sizeof(T) == 8 ? ncclFloat64 : ncclFloat32;
    for (size_t j = 0; j < h->send_peers.size(); ++j) {
        const ptrdiff_t total = h->send_count[j] * h->bs;
        if (total == 0) continue;
        const int grid = (int)std::min<ptrdiff_t>((total + kHaloThreads - 1) / kHaloThreads, (ptrdiff_t)sm_count() * 8);
        halo_pack_kernel<T><<<grid, kHaloThreads, 0, s>>>(h->send_count[j], h->send_idx[j].ptr, h->bs, x, (T *)h->send_buf[j]);
        B2_CHECK_LAUNCH();
    }
    NcclApi *api = nccl_api();
    B2_NCCL(api->GroupStart());
    for (size_t j = 0; j < h->send_peers.size(); ++j)
        if (h->send_count[j]) B2_NCCL(api->Send(h->send_buf[j], (size_t)(h->send_count[j] * h->bs), dt, h->send_peers[j], c->nccl, s));
    const bool indexed = !h->recv_idx.empty();
    for (size_t j = 0; j < h->recv_peers.size(); ++j)
        if (h->recv_count[j])
            B2_NCCL(api->Recv(indexed ? h->recv_stage[j] : (void *)(x + h->recv_begin[j] * h->bs), (size_t)(h->recv_count[j] * h->bs), dt,
                              h->recv_peers[j], c->nccl, s));
    B2_NCCL(api->GroupEnd());
    if (indexed) {
        for (size_t j = 0; j < h->recv_peers.size(); ++j) {
            const ptrdiff_t total = h->recv_count[j] * h->bs;
            if (total == 0) continue;
            const int grid = (int)std::min<ptrdiff_t>((total + kHaloThreads - 1) / kHaloThreads, (ptrdiff_t)sm_count() * 8);
            halo_unpack_kernel<T><<<grid, kHaloThreads, 0, s>>>(h->recv_count[j], h->recv_idx[j].ptr, h->bs, (const T *)h->recv_stage[j], x);
            B2_CHECK_LAUNCH();
        }
    }
    return B200FEM_SUCCESS;
}

}  // namespace b200fem

extern "C" {

int b200fem_comm_unique_id(void *id_out) {
    if (!id_out) B2_FAIL("comm_unique_id: NULL output");
    NcclApi *api = nccl_api();
    if (!api->ok) B2_FAIL(api->why);
    ncclUniqueId id;
    B2_NCCL(api->GetUniqueId(&id));
    memcpy(id_out, &id, B200FEM_UNIQUE_ID_BYTES);
    return B200FEM_SUCCESS;
}

int b200fem_comm_create(b200fem_comm_t *out, int world, int rank, const void *unique_id) {
    if (!out || world < 1 || rank < 0 || rank >= world) B2_FAIL("comm_create: bad arguments");
    auto c = new b200fem_comm_s();
    c->world = world;
    c->rank = rank;
    if (cudaGetDevice(&c->device) != cudaSuccess) {
        delete c;
        B2_FAIL("comm_create: no CUDA device");
    }
    if (cudaMalloc((void **)&c->dev_scratch, sizeof(double) * b200fem_comm_s::kScratch) != cudaSuccess ||
        cudaMallocHost((void **)&c->host_scratch, sizeof(double) * b200fem_comm_s::kScratch) != cudaSuccess ||
        cudaHostAlloc((void **)&c->err_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&c->err_dev, c->err_host, 0) != cudaSuccess) {
        cuda_ok(cudaGetLastError(), "comm_create allocations", __FILE__, __LINE__);
        b200fem_comm_destroy(c);
        return B200FEM_FAILURE;
    }
    *c->err_host = 0;
    if (const char *t = getenv("B200FEM_COMM_TIMEOUT_S")) {
        const double sec = atof(t);
        if (sec > 0) c->timeout_ns = (unsigned long long)(sec * 1e9);
    }
    if (world > 1) {
        NcclApi *api = nccl_api();
        if (!api->ok || !unique_id) {
            set_error(api->ok ? "comm_create: NULL unique id" : api->why);
            b200fem_comm_destroy(c);
            return B200FEM_FAILURE;
        }
        ncclUniqueId id;
        memcpy(&id, unique_id, B200FEM_UNIQUE_ID_BYTES);
        ncclResult_t r = api->CommInitRank(&c->nccl, world, id, rank);
        if (r != ncclSuccess) {
            set_error(std::string("ncclCommInitRank: ") + api->GetErrorString(r));
            c->nccl = nullptr;
            b200fem_comm_destroy(c);
            return B200FEM_FAILURE;
        }
    }
    *out = c;
    return B200FEM_SUCCESS;
}

int b200fem_comm_destroy(b200fem_comm_t c) {
    if (!c) return B200FEM_SUCCESS;
    if (c->nccl) nccl_api()->CommDestroy(c->nccl);
    if (c->dev_scratch) cudaFree(c->dev_scratch);
    if (c->host_scratch) cudaFreeHost(c->host_scratch);
    if (c->err_host) cudaFreeHost(c->err_host);
    delete c;
    return B200FEM_SUCCESS;
}

int b200fem_comm_rank(b200fem_comm_t c) { return c ? c->rank : 0; }
int b200fem_comm_size(b200fem_comm_t c) { return c ? c->world : 1; }

int b200fem_comm_error(b200fem_comm_t c) {
    if (!c) return 0;
    return *((volatile int *)c->err_host);
}

int b200fem_comm_allreduce_sum(b200fem_comm_t c, double *values_dev, int n, void *stream) {
    if (!c || c->world == 1 || n <= 0) return B200FEM_SUCCESS;
    B2_NCCL(nccl_api()->AllReduce(values_dev, values_dev, (size_t)n, ncclFloat64, ncclSum, c->nccl, as_stream(stream)));
    return B200FEM_SUCCESS;
}

int b200fem_comm_sum_to_host(b200fem_comm_t c, const double *values_dev, int n, double *values_host, void *stream) {
    if (!c) B2_FAIL("comm_sum_to_host: NULL communicator");
    if (n <= 0) return B200FEM_SUCCESS;
    if (n > b200fem_comm_s::kScratch) B2_FAIL("comm_sum_to_host: at most 64 values per call");
    cudaStream_t s = as_stream(stream);
    const double *src = values_dev;
    if (c->world > 1) {
        B2_NCCL(nccl_api()->AllReduce(values_dev, c->dev_scratch, (size_t)n, ncclFloat64, ncclSum, c->nccl, s));
        src = c->dev_scratch;
    }
    B2_CUDA(cudaMemcpyAsync(c->host_scratch, src, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) values_host[i] = c->host_scratch[i];
    if (*((volatile int *)c->err_host)) B2_FAIL("halo exchange timed out waiting for a peer (B200FEM_COMM_TIMEOUT_S)");
    return B200FEM_SUCCESS;
}

int b200fem_comm_allreduce_host(b200fem_comm_t c, double *values_host, int n, int op, void *stream) {
    if (!c) B2_FAIL("comm_allreduce_host: NULL communicator");
    if (n <= 0 || c->world == 1) return B200FEM_SUCCESS;
    if (n > b200fem_comm_s::kScratch) B2_FAIL("comm_allreduce_host: at most 64 values per call");
    cudaStream_t s = as_stream(stream);
    for (int i = 0; i < n; ++i) c->host_scratch[i] = values_host[i];
    B2_CUDA(cudaMemcpyAsync(c->dev_scratch, c->host_scratch, sizeof(double) * n, cudaMemcpyHostToDevice, s));
    B2_NCCL(nccl_api()->AllReduce(c->dev_scratch, c->dev_scratch, (size_t)n, ncclFloat64, op == B200FEM_REDUCE_MAX ? ncclMax : ncclSum,
                                  c->nccl, s));
    B2_CUDA(cudaMemcpyAsync(c->host_scratch, c->dev_scratch, sizeof(double) * n, cudaMemcpyDeviceToHost, s));
    B2_CUDA(cudaStreamSynchronize(s));
    for (int i = 0; i < n; ++i) values_host[i] = c->host_scratch[i];
    return B200FEM_SUCCESS;
}

int b200fem_comm_barrier(b200fem_comm_t c, void *stream) {
    double v = 0;
    return b200fem_comm_allreduce_host(c, &v, 1, B200FEM_REDUCE_SUM, stream);
}

int b200fem_halo_create(b200fem_halo_t *out, b200fem_comm_t comm, int real_type, int block_size, ptrdiff_t n_owned,
                        ptrdiff_t n_local, int n_send_peers, const int *send_peers, const ptrdiff_t *send_offsets,
                        const int32_t *send_idx, int n_recv_peers, const int *recv_peers, const ptrdiff_t *recv_begin,
                        const ptrdiff_t *recv_count, int transport) {
    if (!out || !comm) B2_FAIL("halo_create: NULL argument");
    if (real_type != B200FEM_F64 && real_type != B200FEM_F32) B2_FAIL("halo_create: bad real_type");
    if (block_size < 1 || n_owned < 0 || n_local < n_owned) B2_FAIL("halo_create: bad sizes");
    auto h = new b200fem_halo_s();
    h->comm = comm;
    h->real_type = real_type;
    h->bs = block_size;
    h->n_owned = n_owned;
    h->n_local = n_local;
    for (int j = 0; j < n_send_peers; ++j) {
        h->send_peers.push_back(send_peers[j]);
        h->send_count.push_back(send_offsets[j + 1] - send_offsets[j]);
    }
    h->send_idx = std::vector<DeviceArray<int32_t>>((size_t)n_send_peers);
    for (int j = 0; j < n_send_peers; ++j) {
        for (ptrdiff_t i = send_offsets[j]; i < send_offsets[j + 1]; ++i)
            if (send_idx[i] < 0 || send_idx[i] >= n_owned) {
                halo_release(h);
                B2_FAIL("halo_create: a send index is not an owned node");
            }
        if (h->send_idx[(size_t)j].upload(send_idx + send_offsets[j], (size_t)h->send_count[(size_t)j])) {
            halo_release(h);
            return B200FEM_FAILURE;
        }
    }
    for (int j = 0; j < n_recv_peers; ++j) {
        if (recv_begin[j] < n_owned || recv_begin[j] + recv_count[j] > n_local) {
            halo_release(h);
            B2_FAIL("halo_create: a receive range is outside the ghost block");
        }
        h->recv_peers.push_back(recv_peers[j]);
        h->recv_begin.push_back(recv_begin[j]);
        h->recv_count.push_back(recv_count[j]);
    }
    h->transport = B200FEM_HALO_NCCL;
    if (comm->world > 1) {
        if (transport == B200FEM_HALO_AUTO || transport == B200FEM_HALO_PUSH) {
            if (halo_setup_push(h) == B200FEM_SUCCESS) {
                h->transport = B200FEM_HALO_PUSH;
            } else if (transport == B200FEM_HALO_PUSH) {
                halo_release(h);
                return B200FEM_FAILURE;
            }
        }
        if (h->transport == B200FEM_HALO_NCCL) {
            for (int j = 0; j < n_send_peers; ++j) {
                void *p = nullptr;
                const size_t bytes = std::max<size_t>(256, (size_t)h->send_count[(size_t)j] * block_size * real_type);
                if (cudaMalloc(&p, bytes) != cudaSuccess) {
                    cuda_ok(cudaGetLastError(), "halo send buffer", __FILE__, __LINE__);
                    halo_release(h);
                    return B200FEM_FAILURE;
                }
                h->send_buf.push_back(p);
            }
        }
        // nobody starts pushing before every rank's flags and buffers exist
        if (b200fem_comm_barrier(comm, nullptr)) {
            halo_release(h);
            return B200FEM_FAILURE;
        }
    }
    *out = h;
    return B200FEM_SUCCESS;
}

int b200fem_halo_destroy(b200fem_halo_t h) {
    if (h && h->comm && h->comm->world > 1) {
        // peers may still be reading my receive block / writing my flags: drain, then agree that everybody is done
        cudaDeviceSynchronize();
        b200fem_comm_barrier(h->comm, nullptr);
    }
    halo_release(h);
    return B200FEM_SUCCESS;
}

int b200fem_halo_set_recv_index(b200fem_halo_t h, const ptrdiff_t *recv_offsets, const int32_t *recv_idx) {
    if (!h || !recv_offsets) B2_FAIL("halo_set_recv_index: NULL argument");
    const size_t np = h->recv_peers.size();
    std::vector<DeviceArray<int32_t>> idx(np);
    for (size_t j = 0; j < np; ++j) {
        const ptrdiff_t n = recv_offsets[j + 1] - recv_offsets[j];
        if (n != h->recv_count[j]) B2_FAIL("halo_set_recv_index: count differs from the plan's receive count");
        for (ptrdiff_t i = recv_offsets[j]; i < recv_offsets[j + 1]; ++i)
            if (recv_idx[i] < h->n_owned || recv_idx[i] >= h->n_local) B2_FAIL("halo_set_recv_index: an index is not a ghost node");
        if (idx[j].upload(recv_idx + recv_offsets[j], (size_t)n)) return B200FEM_FAILURE;
    }
    h->recv_idx = std::move(idx);
    if (h->transport == B200FEM_HALO_PUSH) {
        for (size_t j = 0; j < np; ++j) h->recv.p[j].idx = h->recv_idx[j].ptr;
    } else if (h->comm->world > 1) {
        h->recv_stage.assign(np, nullptr);
        for (size_t j = 0; j < np; ++j)
            if (h->recv_count[j]) B2_CUDA(cudaMalloc(&h->recv_stage[j], (size_t)h->recv_count[j] * (size_t)h->bs * (size_t)h->real_type));
    }
    return B200FEM_SUCCESS;
}
int b200fem_halo_transport(b200fem_halo_t h) { return h ? h->transport : -1; }
ptrdiff_t b200fem_halo_values_sent(b200fem_halo_t h) {
    ptrdiff_t n = 0;
    if (h)
        for (auto v : h->send_count) n += v * h->bs;
    return n;
}
int b200fem_halo_launches_per_gather(b200fem_halo_t h) {
    if (!h || h->comm->world == 1) return 0;
    if (h->transport == B200FEM_HALO_PUSH) return (h->send_peers.empty() ? 0 : 1) + (h->recv_peers.empty() ? 0 : 1);
    return (int)h->send_peers.size();
}

int b200fem_halo_gather(b200fem_halo_t h, void *x_local, void *stream) {
    if (!h || !x_local) B2_FAIL("halo_gather: NULL argument");
    if (h->real_type == B200FEM_F64) return halo_gather_t<double>(h, (double *)x_local, as_stream(stream));
    return halo_gather_t<float>(h, (float *)x_local, as_stream(stream));
}
int b200fem_halo_gather_begin(b200fem_halo_t h, void *x_local, void *stream) {
    if (!h || !x_local) B2_FAIL("halo_gather_begin: NULL argument");
    if (h->real_type == B200FEM_F64) return halo_gather_t<double>(h, (double *)x_local, as_stream(stream), 1);
    return halo_gather_t<float>(h, (float *)x_local, as_stream(stream), 1);
}
int b200fem_halo_gather_deferred_begin(b200fem_halo_t h, void *x_local, void *stream, const unsigned long long **ready_flag,
                                       unsigned long long *ready_value) {
    if (!h || !x_local || !ready_flag || !ready_value) B2_FAIL("halo_gather_deferred_begin: NULL argument");
    *ready_flag = nullptr;
    *ready_value = 0;
    if (h->comm->world == 1 || h->transport != B200FEM_HALO_PUSH || h->recv_peers.empty()) {
        // nothing to wait for in a kernel: the ordinary gather on `stream`
        return b200fem_halo_gather(h, x_local, stream);
    }
    if (!h->side) {
        // highest priority: the gather's (few, short) blocks must not queue behind the apply's persistent CTAs
        int lo = 0, hi = 0;
        B2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        B2_CUDA(cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, hi));
        B2_CUDA(cudaEventCreateWithFlags(&h->side_event, cudaEventDisableTiming));
        B2_CUDA(cudaMalloc((void **)&h->ready, 2 * sizeof(unsigned long long)));
        B2_CUDA(cudaMemset(h->ready, 0, 2 * sizeof(unsigned long long)));
    }
    // the push reads x, the copy writes its ghost block: behind everything `stream` has been given so far (the producers of
    // x, the previous apply that read the block), concurrent with what follows on `stream`
    B2_CUDA(cudaEventRecord(h->side_event, as_stream(stream)));
    B2_CUDA(cudaStreamWaitEvent(h->side, h->side_event, 0));
    *ready_flag = h->ready;
    *ready_value = ++h->epoch;
    return B200FEM_SUCCESS;
}
int b200fem_halo_gather_deferred_launch(b200fem_halo_t h, void *x_local) {
    if (!h || !x_local) B2_FAIL("halo_gather_deferred_launch: NULL argument");
    if (!h->side) B2_FAIL("halo_gather_deferred_launch: no deferred gather has been begun");
    return h->real_type == B200FEM_F64 ? halo_gather_t<double>(h, (double *)x_local, nullptr, 3) : halo_gather_t<float>(h, (float *)x_local, nullptr, 3);
}
int b200fem_halo_gather_end(b200fem_halo_t h, void *x_local, void *stream) {
    if (!h || !x_local) B2_FAIL("halo_gather_end: NULL argument");
    if (h->real_type == B200FEM_F64) return halo_gather_t<double>(h, (double *)x_local, as_stream(stream), 2);
    return halo_gather_t<float>(h, (float *)x_local, as_stream(stream), 2);
}

}  // extern "C"
