// One-shot exchange of a rank's [sum of target scores, #foreground] over peer-mapped mailboxes (sm_100a, NVLink 5 /
// NVSwitch): the task-aligned path's one real exchange step, fused into the kernels on both sides
// instead of a collective call between them.
//
//   producer   tal_stats_kernel (csrc/tal.cu), the last kernel of yb_tal_assign: once the rank's statistics are summed,
//              `world` of its threads store them straight into every rank's mailbox (peer stores over NVLink; the own
//              mailbox is just one of the targets).
//   consumer   peer_wait_kernel, the first kernel of yb_tal_loss: one thread per rank polls the OWN (local) mailbox
//              until that rank's entry of this step has arrived, then the entries are averaged in rank order — every
//              rank adds the same numbers in the same order, so all ranks use the bit-identical normaliser.
//
// An entry is two 8-byte words, each carrying the step's sequence number next to its payload (single-copy atomic: no
// fence, no flag/payload ordering to get wrong); entries live in a ring of kPeerSlots steps, so a rank that runs ahead
// (by at most one step: its next exchange needs this one's result) never overwrites what a slower rank still reads.
// NCCL needs ~35 us for this 8-byte message (launch + protocol); this costs a store and a poll.
//
// Set-up (once): every rank allocates its mailbox with yb_peer_mailbox_alloc, exports a CUDA IPC handle, the handles are
// exchanged by whatever the host has (torch.distributed.all_gather_object in the Python binding) and opened with
// yb_peer_mailbox_open.  Single node only: the ranks must be able to map each other's memory.
#include <cstring>

#include "common.cuh"

namespace yb {

constexpr int kPeerSlots = 4;

__global__ void __launch_bounds__(32)
peer_wait_kernel(const unsigned long long *__restrict__ mailbox, int world, unsigned int seq, float *__restrict__ out2,
                 unsigned int *__restrict__ timed_out) {
    __shared__ float s_t[YB_PEER_MAX_WORLD], s_n[YB_PEER_MAX_WORLD];
    const int r = threadIdx.x;
    if (r < world) {
        const volatile unsigned long long *e = mailbox + ((size_t)(seq % kPeerSlots) * YB_PEER_MAX_WORLD + r) * 2;
        unsigned long long w0 = 0, w1 = 0;
        const long long t0 = clock64();
        bool ok = false;
        for (;;) {
            w0 = e[0]; w1 = e[1];
            ok = (unsigned int)(w0 >> 32) == seq && (unsigned int)(w1 >> 32) == seq;
            if (ok || clock64() - t0 > (20ll << 30)) break;          // ~10 s at 2 GHz: a peer is gone; do not hang the GPU
            __nanosleep(64);
        }
        if (!ok) atomicOr(timed_out, 1u);
        s_t[r] = ok ? __uint_as_float((unsigned int)w0) : __int_as_float(0x7fc00000);
        s_n[r] = ok ? __uint_as_float((unsigned int)w1) : __int_as_float(0x7fc00000);
    }
    __syncwarp();
    if (r == 0) {
        double t = 0.0, n = 0.0;
        for (int k = 0; k < world; ++k) { t += (double)s_t[k]; n += (double)s_n[k]; }   // rank order: identical on every rank
        out2[0] = (float)(t / (double)world);
        out2[1] = (float)(n / (double)world);
    }
}

int launch_peer_wait(const yb_peer_exchange &px, float *out2, unsigned int *timed_out, cudaStream_t st) {
    peer_wait_kernel<<<1, 32, 0, st>>>(static_cast<const unsigned long long *>(px.mailbox[px.rank]), px.world, px.seq, out2,
                                       timed_out);
    YB_LAUNCH_CHECK();
    return YB_OK;
}

int check_peer(const yb_peer_exchange *px, const char *who) {
    YB_REQUIRE(px->world >= 1 && px->world <= YB_PEER_MAX_WORLD, "%s: peer exchange world must be in [1, %d]", who, YB_PEER_MAX_WORLD);
    YB_REQUIRE(px->rank >= 0 && px->rank < px->world, "%s: peer exchange rank %d out of range", who, px->rank);
    YB_REQUIRE(px->seq != 0u, "%s: peer exchange sequence numbers start at 1", who);
    for (int r = 0; r < px->world; ++r) YB_REQUIRE(px->mailbox[r] != nullptr, "%s: mailbox of rank %d is not mapped", who, r);
    return YB_OK;
}

}  // namespace yb

using namespace yb;

extern "C" size_t yb_peer_mailbox_bytes(void) { return sizeof(unsigned long long) * 2 * YB_PEER_MAX_WORLD * kPeerSlots; }

extern "C" int yb_peer_mailbox_alloc(void **mailbox_out) {
    YB_REQUIRE(mailbox_out != nullptr, "yb_peer_mailbox_alloc: null pointer");
    void *p = nullptr;
    YB_CUDA(cudaMalloc(&p, yb_peer_mailbox_bytes()));          // plain cudaMalloc: the one kind of memory CUDA IPC can export
    YB_CUDA(cudaMemset(p, 0, yb_peer_mailbox_bytes()));
    YB_CUDA(cudaDeviceSynchronize());
    *mailbox_out = p;
    return YB_OK;
}

extern "C" int yb_peer_mailbox_free(void *mailbox) {
    if (mailbox != nullptr) YB_CUDA(cudaFree(mailbox));
    return YB_OK;
}

extern "C" int yb_peer_mailbox_export(void *mailbox, void *handle_out_64) {
    YB_REQUIRE(mailbox && handle_out_64, "yb_peer_mailbox_export: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == YB_PEER_HANDLE_BYTES, "CUDA IPC handle size");
    cudaIpcMemHandle_t h;
    YB_CUDA(cudaIpcGetMemHandle(&h, mailbox));
    memcpy(handle_out_64, &h, sizeof(h));
    return YB_OK;
}

extern "C" int yb_peer_mailbox_open(const void *handle_64, void **peer_mailbox_out) {
    YB_REQUIRE(handle_64 && peer_mailbox_out, "yb_peer_mailbox_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, sizeof(h));
    void *p = nullptr;
    YB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *peer_mailbox_out = p;
    return YB_OK;
}

extern "C" int yb_peer_mailbox_close(void *peer_mailbox) {
    if (peer_mailbox != nullptr) YB_CUDA(cudaIpcCloseMemHandle(peer_mailbox));
    return YB_OK;
}
