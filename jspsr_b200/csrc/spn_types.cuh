// Halo-independent types shared by abi.cu and both kernel variants.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace jspsr {

constexpr int TILE_W = 128;   // columns per CTA
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

enum { NORM_NONE = 0, NORM_RESIDUAL = 1, NORM_SUM = 2 };

// Geometry of one call.  Rows are expressed in GLOBAL image coordinates so that
// a row strip (multi-GPU sharding) computes bit-identical positions.
struct Geom {
    int B, H, W;          // rows/cols of weight/offset/out held by this call (the strip)
    int H_img;            // rows of the whole image
    int row0;             // global row of out row 0
    int init_row0;        // global row of init buffer row 0
    int init_rows;        // rows present in the init buffer
    int tiles_x, tiles_y;
};

// Row-strip halo exchange fused into the forward (include/jspsr_peer.h, jspsr_strip_peer): device-side view.
// Flags are monotonically increasing generation stamps; `wait_*` are this rank's own flag words (raised by the
// neighbours), `*_flag` / `*_dst` live in the neighbours' memory (CUDA IPC mappings reached over NVLink).
struct StripPeerDev {
    void* up_dst = nullptr;             // upper neighbour's bottom halo rows of its next DEM buffer (or null: signal only)
    void* dn_dst = nullptr;             // lower neighbour's top halo rows
    unsigned* up_flag = nullptr;        // raised to stamp + 1 when this strip's top edge CTAs are done (null: first strip)
    unsigned* dn_flag = nullptr;
    const unsigned* wait_up = nullptr;  // raised by the upper neighbour when its rows of generation `stamp` have landed here
    const unsigned* wait_dn = nullptr;
    unsigned* tickets = nullptr;        // [2] local counters of finished edge CTAs (zero between launches)
    unsigned stamp = 0;
    int halo = 0;
    int debug = 0;  // JSPSR_STRIP_PEER_DEBUG (timing experiments only, results are NOT valid): 1 = plain tile order,
                    // 2 = no row copy in the push, 4 = no waits, 8 = no push at all (with 4)
};

// Gradient all-reduce fused into the backward (jspsr_peer_reduce): slots[p] = rank p's buffer as mapped here,
// laid out [2 parities][8 source ranks][16 doubles] followed by the device-side step counter.
struct PeerReduceDev {
    double* slots[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int rank = 0, world = 1;
    float mul = 1.f;  // 1 / world when averaging
};
constexpr int REDUCE_SLOT = 16;                       // doubles per (parity, source) slot; [15] holds the stamp
constexpr int REDUCE_COUNTER_OFFSET = 2 * 8 * REDUCE_SLOT;  // in doubles: the step counter follows the slots

// Host-side launch descriptor (filled by abi.cu)
struct LaunchArgs {
    const void* init = nullptr;
    const void* weight = nullptr;
    const void* offset = nullptr;
    const float* w9 = nullptr;
    const float* b1 = nullptr;
    void* out = nullptr;
    // backward only
    const void* grad_out = nullptr;
    float* grad_init = nullptr;
    void* grad_weight = nullptr;
    void* grad_offset = nullptr;
    float* grad_w9 = nullptr;
    float* grad_b1 = nullptr;
    void* workspace = nullptr;
    bool accumulate = false;
    bool gen_preact = false;  // JSPSR_BWD_GEN_PREACT: grad_weight is [B,25,H,W] pre-activation gradients
    Geom g{};
    int mode = NORM_RESIDUAL;
    float scale = 1.f;
    bool bf16 = false;      // weight / offset (and their gradients) are bf16
    bool init_f32 = false;  // with bf16: init / out / grad_out stay fp32 (JSPSR_MIXED, what torch.autocast produces)
    bool use_tma = false;
    bool pair = false;      // bf16 weight / offset streamed as bf16x2 words, two pixels per thread (abi.cu: pair_ok)
    int* status = nullptr;
    int tile_h = 16;  // rows per CTA (16 / 8 / 4 / 2), chosen by abi.cu; the TMA box is encoded to match
    cudaStream_t stream = nullptr;
    CUtensorMap tmap{};
    const StripPeerDev* strip_peer = nullptr;    // forward: fused halo exchange (B = 1, 16 rows per CTA)
    const PeerReduceDev* peer_reduce = nullptr;  // backward: fused all-reduce of grad_w9 / grad_b1
};


// reduction workspace layout (caller-owned, zero on entry, zero on exit)
struct alignas(16) ReduceWs {
    double sums[12];          // grad_w[0..8], grad_b, spare
    unsigned int ticket;      // CTAs that have contributed
    unsigned int pad[3];
};


// Opt a kernel in to more than 48 KB of shared memory, once per (kernel, device): the attribute is per device, and
// a process may drive several devices.  Host only; not a stream operation (safe while a stream is being captured).
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);

}  // namespace jspsr
