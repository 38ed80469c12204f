// Internal declarations shared by the translation units of libies_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/ies_b200.h"

namespace ies {

constexpr int MAX_PAL = 32;     // palette entries carried in the kernel parameters
constexpr int MAX_TERMS = 12;   // 6 faces x 2 components per half-step (space.py:1054-1108)

struct Box { int lo[3], hi[3]; };

struct PmlTermDev {
    int comp, diff, axis;
    int lo[3], hi[3];
    int psi_off;
    int pdim[3];                // psi array extents
    double sign;
    void* psi;                  // field dtype
    const double *b, *a, *kf;   // per-cell-along-axis tables (device)
};

// Everything one half-step update kernel needs, passed by value.
struct UpdParams {
    const void* F[3];           // differentiated field (E in updateH, H in updateE): x,y,z
    void* G[3];                 // updated field
    const double* C;            // CH2 / CE2 (space.py:445-553, zero conductivity)
    const uint8_t* Cidx;        // lossless palette form of C (<= MAX_PAL distinct values) or null:
    double cpal[MAX_PAL];       //   C[i] == cpal[Cidx[i]] bit for bit; 1 B/cell of HBM traffic instead of 8,
                                //   the palette itself sits in the kernel-parameter constant bank
    const double* Ctile;        // per y-line tile (plane, column block): the tile's one coefficient value, NaN if not uniform; or null
    const void* halo[2];        // neighbour planes of F_y, F_z (or null)
    const void* dz[2];          // scratch: d/dz F_y, d/dz F_x   (spectral methods)
    const void* dxs[2];         // scratch: d/dx F_z, d/dx F_y   (PSTD)
    const void* dys[2];         // scratch: d/dy F_z, d/dy F_x   (direct-circulant path only)
    int nx, ny, nz;
    int dir;                    // +1: forward differences (updateH); -1: backward (updateE)
    int i0, i1;                 // x range handled by this launch
    int pstd;                   // x derivative comes from dxs[]
    int dz_discard;             // fused kernel: drop the consumed dz scratch lines from L2 (discard.global.L2), no write-back
    long long dz_off;           // element offset of the dz scratch relative to the field index
    double rdx, rdy, rdz;
    Box box[3];
    // y-derivative side buffer of the separate CPML pass (engine.cu k_pml_terms): the update phase of the
    // spectral kernels saves (d/dy F_z, d/dy F_x) of the rows [0, ys_lo_n) and [ys_hi_0, ny) there,
    // index ((f * nx + i) * ys_rows + jj) * nz + k with jj = j or j - ys_hi_0 + ys_lo_n; null = nothing to save
    void* dy_side;
    int ys_lo_n, ys_hi_0, ys_rows;
    int fdtd;                   // derivatives are two-point differences (k_pml_terms recomputes them)
    int nterms;
    short dz_keep_lo, dz_keep_hi;   // dz_discard: only lines with columns inside [dz_keep_lo, dz_keep_hi) are dropped (the CPML pass reads the rest)
    PmlTermDev terms[MAX_TERMS];
};

struct Ctx {
    ies_config cfg;
    int esize;                  // bytes per field element
    bool cplx, dbl;
    cudaStream_t stream, own_stream;
    void* F[6];
    double* C[2];
    uint8_t* Cidx[2];           // palette-compressed coefficients (valid when Cnpal > 0)
    double* Cpal[2];            // device scratch of the palette builder (256 slots + overflow flag)
    double Cpal_host[2][MAX_PAL];
    int Cnpal[2];
    int use_palette;
    double* Ctile[2];           // per-tile uniform coefficient of the y-line kernel's tiles (or null)
    int use_ctile;
    int fdtd_vec;               // FDTD: 16-byte vectorised kernel when nz allows (k_fdtd_vec)
    void* scratch[6];           // dzA dzB dxA dxB (+ dyA dyB on the direct-circulant path)
    bool generic;               // a spectral axis is not a power of two in 16..512: derivatives by direct circulant sums
    void* circ[2][3];           // [half][axis] first column of the derivative circulant (FFT-precision complex, n entries)
    void* halo_recv[2][2];      // [half][y|z] planes inside halo_block
    void* halo_block;           // one cudaMalloc: 4 recv planes + the arrival flags (one IPC handle)
    size_t halo_block_bytes;
    unsigned* halo_flag[2];     // [half] push counter written by the neighbour (inside halo_block)
    // inter-process neighbours (CUDA IPC): mapped halo blocks of rank-1 (0) and rank+1 (1)
    void* peer_block[2];
    void* peer_recv[2][2][2];   // [nbr][half][y|z] recv planes of that neighbour
    unsigned* peer_flag[2][2];  // [nbr][half]
    unsigned push_seq[2], wait_seq[2];
    unsigned* seq_ring;         // pinned host ring of sequence numbers (memcpy fallback of the flag write)
    int flag_write_mode;        // 0 = cuStreamWriteValue32, 1 = cudaMemcpyAsync from seq_ring
    void* mult[2][3];           // [half][axis] complex table in FFT precision, pre-scaled by 1/N
    void* tw[3];                // W_N master twiddles per axis
    void* tw_t[3];              // stage tables transposed to [m][jj], [forward | inverse] (fft_dev.cuh TwTables)
    Box ubox[6];
    std::vector<PmlTermDev> terms[2];
    std::vector<void*> owned;   // device allocations freed at destroy
    int ghost_on[3];
    double ghost_pp[3][2], ghost_pm[3][2];
    int has_prev, has_next;
    void* stage; size_t stage_bytes;   // staging buffer for get/set
    void* src_tab; size_t src_tab_bytes;          // Setter phase tables px|py|pz (ies_put_src), device copy
    std::vector<double> src_tab_host;             //   and the host shadow it was uploaded from
    cudaEvent_t ev_halo;
    cudaEvent_t ev_t0, ev_t1;          // ies_timer_start/stop
    int profiling;                     // per-kernel CUDA-event timing on/off
    std::vector<cudaEvent_t> prof_ev[4][2];   // [slot][begin/end]
    // fused single-launch SHPF half-step (shpf_fused.cuh)
    int use_pml_split;                 // CPML corrections: 1 = separate pass over the absorber cells (k_pml_terms),
                                       // 0 = inside the update kernels, -1 (default) = separate when a y or z face has terms
    void* dy_side; size_t dy_side_bytes;
    int use_fused;                     // 1 = k_shpf_fused where instantiated, 0 = k_zline + k_yline_update
    int fused_prefetch;                // z role prefetches F_z and G of its rows into L2 for the y role
    int fused_discard;                 // y role discards the dz scratch lines it consumed (no DRAM write-back of the scratch)
    int fused_zb;                      // z tiles per z-role CTA (1, 2 or 4)
    int fused_lead, fused_ring_planes; // planes of lead of the z role; scratch ring size in planes
    unsigned* fused_sync;              // ticket + zdone[nx] + ydone[nx]
    void* fused_ring[2];               // ring scratch (fused_ring_planes planes each) or null
    int fused_ring_alloc;              // planes the ring buffers were allocated for
    unsigned long long* fused_prof_mem;
    unsigned long long* fused_prof;    // development: per-phase cycle counters of the fused kernel (option fused_prof) or null
    void* twz_t;                       // z axis: transposed stage tables [forward | inverse] (fft_dev.cuh TwTables)
};

enum { PROF_ZLINE = 0, PROF_YLINE_UPDATE = 1, PROF_XLINE = 2, PROF_FDTD = 3 };
void prof_mark(Ctx* c, int slot, int end);

void set_error(const std::string& s);
void count_launch(int n = 1);
#define IES_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    ies::set_error(std::string(#x) + ": " + cudaGetErrorString(e_)); return 1; } } while (0)

// spectral_*.cu: z-line / strided-line derivative passes and the fused y-line update.
// All return 0 or set the error and return 1.
template <typename T, bool CPLX>
int launch_zline(Ctx* c, const void* A, const void* B, void* dA, void* dB, int half, int i0, int i1, int out_i0,
                 cudaStream_t st = nullptr);
template <typename T, bool CPLX>
int launch_sline(Ctx* c, const void* A, const void* B, void* dA, void* dB, int half, int axis, int i0, int i1);
template <typename T, bool CPLX>
int launch_yline_update(Ctx* c, const UpdParams& p, int half);
template <typename T, bool CPLX>
int launch_shpf_fused(Ctx* c, const UpdParams& p, int half);       // shpf_fused.cuh; 2 = not instantiated
int dev_alloc(Ctx* c, void** p, size_t bytes, bool zero = true);

bool fft_len_supported(int n);

}  // namespace ies
