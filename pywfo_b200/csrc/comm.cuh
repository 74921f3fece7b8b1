// K5: the path's one collective, inside the library.
//
// The ranks of ONE node each own a GPU and a process.  Every rank allocates a communication buffer
// in its HBM, exports it with cudaIpcGetMemHandle and maps the peers' buffers (NVLink / NVSwitch
// peer access), so that a kernel on rank r can store straight into rank r' memory.  Two operations:
//
//   all-reduce(sum) of the n_bra x n_ket partial matrices: ONE-SHOT.  A push kernel stores the
//     partial into slot [rank] of every peer's buffer and raises a flag there; a second kernel waits
//     for all flags and adds the slots in rank order -- every rank adds the same numbers in the same
//     order, so all ranks hold bitwise identical results (NCCL's ring does not promise that), and
//     the 20 KB exchange costs one NVLink store latency instead of a host-launched library call;
//   all-gather of the ket CI tensor: every rank uploads 1/world of it from the host and pushes that
//     slice to all peers (8 x 22 MB over PCIe at the same time was the e2e bottleneck in round 1).
//
// Flags are epoch counters (never reset); the slots are double buffered by epoch parity: a peer can
// be at most one collective ahead of the slowest rank, because finishing collective k needs every
// rank's push of k.  A wait that does not complete in ~20 s (a rank died) sets an error word instead
// of spinning forever.
//
// Bootstrap: the 64-byte IPC handles are exchanged through files in /dev/shm named by a job tag the
// caller supplies (wfo_comm_init); no MPI, no torch in the library.
#pragma once
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <thread>

#include "common.cuh"

namespace wfo {

constexpr int COMM_MAX_RANKS = 16;
constexpr size_t COMM_HDR_BYTES = 4096;              // flags + error word
constexpr size_t COMM_RED_CAP = 1u << 20;            // bytes per reduce slot (362 x 362 doubles)
constexpr long long COMM_SPIN_LIMIT = 40000000000ll;  // ~20 s of SM clocks

struct CommHdr {
    int red_flag[COMM_MAX_RANKS];   // written by peer r: epoch of its last reduce push
    int gat_flag[COMM_MAX_RANKS];   // written by peer r: epoch of its last gather push
    int error;                      // set by a wait that timed out
    unsigned done_cnt[COMM_MAX_RANKS];   // local: blocks of a push kernel that finished for peer r
};

struct CommPeers {
    char *base[COMM_MAX_RANKS];
};

struct Comm {
    int rank, world;
    char *buf;           // own buffer (device)
    size_t bytes, gat_cap;
    CommPeers peers;     // base[rank] = buf
    int red_epoch, gat_epoch;
    char tag[128];
};

__host__ __device__ inline size_t comm_red_off(int world, int parity, int src) {
    return COMM_HDR_BYTES + ((size_t)parity * world + src) * COMM_RED_CAP;
}
__host__ __device__ inline size_t comm_gat_off(int world) { return COMM_HDR_BYTES + 2 * (size_t)world * COMM_RED_CAP; }

// store `count` doubles from src to [dst_off] of every peer selected by `to_self` (gather: peers only;
// reduce: self too), then raise flag[flag_base + rank] = epoch there.  grid = world * blocks_per_peer.
__global__ void comm_push_kernel(const double *__restrict__ src, size_t count, CommPeers peers, size_t dst_off,
                                 int flag_word, int rank, int world, int epoch, int blocks_per_peer, bool to_self,
                                 CommHdr *own) {
    const int peer = blockIdx.x / blocks_per_peer, b = blockIdx.x % blocks_per_peer;
    if (peer == rank && !to_self) return;
    double *dst = reinterpret_cast<double *>(peers.base[peer] + dst_off);
    for (size_t i = (size_t)b * blockDim.x + threadIdx.x; i < count; i += (size_t)blocks_per_peer * blockDim.x)
        dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned c = atomicAdd(&own->done_cnt[peer], 1u);
        if (c == (unsigned)blocks_per_peer - 1u) {
            own->done_cnt[peer] = 0;
            __threadfence_system();
            volatile int *flag = reinterpret_cast<volatile int *>(peers.base[peer]) + flag_word + rank;
            *flag = epoch;
        }
    }
}

__device__ __forceinline__ bool comm_wait_flag(const volatile int *flag, int epoch) {
    const long long t0 = clock64();
    while (*flag - epoch < 0) {
        if (clock64() - t0 > COMM_SPIN_LIMIT) return false;
    }
    return true;
}

// wait until every rank in [0, world) except `skip` has raised flag[flag_word + r] to epoch
__global__ void comm_wait_kernel(CommHdr *own, int flag_word, int world, int epoch, int skip) {
    const int r = threadIdx.x;
    if (r < world && r != skip) {
        if (!comm_wait_flag(reinterpret_cast<const volatile int *>(own) + flag_word + r, epoch)) own->error = 1;
    }
}

// out[i] = sum_r slot[r][i] in rank order, after all pushes of this epoch have arrived
__global__ void comm_reduce_sum_kernel(char *own_buf, int world, int parity, int epoch, size_t count, double *out) {
    CommHdr *own = reinterpret_cast<CommHdr *>(own_buf);
    __shared__ int ok;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < world) {
        if (!comm_wait_flag(&own->red_flag[threadIdx.x], epoch)) { own->error = 1; ok = 0; }
    }
    __syncthreads();
    if (!ok) return;
    __threadfence_system();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < world; r++)
            s += reinterpret_cast<const volatile double *>(own_buf + comm_red_off(world, parity, r))[i];
        out[i] = s;
    }
}

// ---- bootstrap through /dev/shm -------------------------------------------------------------
struct CommCard {
    cudaIpcMemHandle_t handle;
    int device;
    int pid;
    unsigned long long bytes;
};

inline std::string comm_path(const char *tag, int rank) {
    std::string t(tag);
    for (auto &c : t)
        if (!(isalnum((unsigned char)c) || c == '-' || c == '_')) c = '_';
    return "/dev/shm/wfo_b200_" + t + "_" + std::to_string(rank) + ".ipc";
}

inline bool comm_write_card(const std::string &path, const CommCard &card) {
    const std::string tmp = path + ".tmp";
    int fd = open(tmp.c_str(), O_CREAT | O_TRUNC | O_WRONLY, 0600);
    if (fd < 0) return false;
    const bool ok = write(fd, &card, sizeof(card)) == (ssize_t)sizeof(card);
    close(fd);
    return ok && rename(tmp.c_str(), path.c_str()) == 0;
}

inline bool comm_read_card(const std::string &path, CommCard *card, double timeout_s) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        int fd = open(path.c_str(), O_RDONLY);
        if (fd >= 0) {
            const bool ok = read(fd, card, sizeof(*card)) == (ssize_t)sizeof(*card);
            close(fd);
            if (ok) return true;
        }
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) return false;
        std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
}

}  // namespace wfo
