/*
 * jspsr_peer.h - the multi-GPU side of libjspsr_spn.so's C ABI: peer-mapped device memory
 * (one process per GPU on one NVSwitch node) and the two places where the propagation path
 * has a real exchange step, each FUSED into the kernel that produces the data:
 *
 *   (1) batch-sharded training (BASELINE configs 3 and 4; DDP semantics around
 *       train/train_utils.py:211-219: loss.backward() -> optimizer.step()).  The path's only
 *       cross-rank state is grad_w[9] / grad_b[1] of PostProcessor.w / .b
 *       (models/components/spn.py:88-89).  jspsr_spn_backward_reduce all-reduces them inside the
 *       backward kernel: its last CTA stores the rank's fp64 sums into its slot of every peer's
 *       buffer (st.release.sys over NVLink), waits for the peers' slots and adds them in rank
 *       order - no NCCL kernel, no extra launch, no stream wait, bit-identical on every rank.
 *
 *   (2) row-strip inference on a large raster (BASELINE config 5; the reference's only
 *       whole-raster path, utils/utils.py:1556-1654, runs on one device).
 *       jspsr_spn_forward_strip_peer writes the first / last `halo` rows it produces straight into
 *       the neighbours' next DEM buffer from its epilogue and raises a flag there; only the CTAs
 *       whose taps can reach halo rows wait for the neighbours' flag, the interior of the band is
 *       computed while the boundary rows travel.
 *
 * The reference has no distributed code at all; nothing here replaces
 * a reference symbol other than the single-device calls cited above.
 *
 * Conventions as in jspsr_spn.h (plain C, device pointers, explicit stream, negative status).
 * Peer memory is the one thing this library allocates: cudaMalloc'd, zero-filled, exported as a
 * CUDA IPC handle the host side hands to the other ranks (torch.distributed is only the
 * messenger for the 64-byte handles).
 */
#ifndef JSPSR_PEER_H_
#define JSPSR_PEER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#ifdef __GNUC__
#pragma GCC visibility push(default)
#endif

#define JSPSR_PEER_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */
#define JSPSR_PEER_MAX_RANKS 8     /* one NVSwitch node */

/* Allocate `bytes` of zero-filled device memory on the current device and export it.
 * handle_out receives JSPSR_PEER_HANDLE_BYTES bytes to be sent to the other ranks.
 * Synchronous (cudaMalloc + cudaMemset + device synchronize): set-up time only. */
int jspsr_peer_alloc(size_t bytes, void **dev_ptr, void *handle_out);
/* Map another rank's allocation into this process (cudaIpcOpenMemHandle, peer access enabled lazily). */
int jspsr_peer_open(const void *handle, void **dev_ptr);
int jspsr_peer_close(void *dev_ptr); /* unmap a pointer returned by jspsr_peer_open */
int jspsr_peer_free(void *dev_ptr);  /* free a pointer returned by jspsr_peer_alloc (all peers must have closed it) */

/* ---- (1) gradient all-reduce fused into the propagation backward ------------------------- */

#define JSPSR_REDUCE_SLOT_DOUBLES 16 /* 10 sums + spare, the last word is the step stamp */
/* bytes of peer memory each rank contributes: [2 parities][JSPSR_PEER_MAX_RANKS][16 doubles] + the step counter */
#define JSPSR_REDUCE_BYTES (2 * JSPSR_PEER_MAX_RANKS * JSPSR_REDUCE_SLOT_DOUBLES * 8 + 64)

typedef struct {
    void *slots[JSPSR_PEER_MAX_RANKS]; /* slots[p]: rank p's JSPSR_REDUCE_BYTES buffer as mapped in THIS process
                                          (slots[rank] is the local allocation) */
    int rank, world;                   /* 1 <= world <= JSPSR_PEER_MAX_RANKS */
    int average;                       /* 1: divide by world (DistributedDataParallel's convention), 0: plain sum */
} jspsr_peer_reduce;

/* jspsr_spn_backward with grad_w9 / grad_b1 all-reduced over the ranks of `pr` before they are written.
 * Every rank must make the same sequence of calls (the step stamp is a device-side counter inside the
 * buffer, so a captured CUDA graph can be replayed).  pr == NULL or world == 1: plain jspsr_spn_backward. */
int jspsr_spn_backward_reduce(const void *grad_out, const void *init, const void *weight,
                              const void *offset, const float *w9, float *grad_init,
                              void *grad_weight, void *grad_offset, float *grad_w9, float *grad_b1,
                              void *workspace, int B, int H, int W, int norm_mode, float scale,
                              int dtype, unsigned flags, const jspsr_peer_reduce *pr, void *stream);

/* ---- (2) row strips with the halo exchange fused into the forward ------------------------ */

/* bytes of (local, peer-visible) flag memory per rank: flags[0] is raised by the upper neighbour,
 * flags[1] by the lower one; [2],[3] are the local tickets of the edge CTAs */
#define JSPSR_STRIP_FLAG_BYTES 64

typedef struct {
    void *up_dst;         /* where my FIRST `halo` output rows go: the bottom halo rows of the upper neighbour's
                             next DEM buffer (peer-mapped); NULL: signal only (last application / T = 1) */
    void *dn_dst;         /* where my LAST `halo` output rows go: the top halo rows of the lower neighbour's next buffer */
    unsigned *up_flags;   /* the upper neighbour's flag block (peer-mapped), NULL on the first strip */
    unsigned *dn_flags;   /* the lower neighbour's flag block (peer-mapped), NULL on the last strip */
    unsigned *my_flags;   /* this rank's flag block (local) */
    unsigned stamp;       /* generation of the DEM buffer this call reads; it waits for flags >= stamp
                             and raises the neighbours' flags to stamp + 1 when its edge rows are done */
    int halo;             /* rows exchanged on each side, <= Hs */
} jspsr_strip_peer;

/* jspsr_spn_forward_strip (B = 1) with the exchange fused in, see the header comment. */
int jspsr_spn_forward_strip_peer(const void *init, const void *weight, const void *offset,
                                 const float *w9, const float *b1, void *out, int Hs, int W,
                                 int H_img, int row0, int init_row0, int init_rows, int norm_mode,
                                 float scale, int dtype, int *status, const jspsr_strip_peer *sp,
                                 void *stream);

/* First generation of a sequence: copy the band's own first / last `halo` rows (band = [Hs,W] rows of the
 * buffer that will be read with sp->stamp) into the neighbours' halo rows (sp->up_dst / sp->dn_dst) and raise
 * their flags to sp->stamp.  Waits for this rank's flags >= stamp - 1 first (the neighbours have finished
 * reading the buffer being overwritten). */
int jspsr_strip_halo_push(const void *band, int Hs, int W, int dtype, const jspsr_strip_peer *sp,
                          void *stream);

#ifdef __GNUC__
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* JSPSR_PEER_H_ */
