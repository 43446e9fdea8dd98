// gevb_internal.cuh -- shared definitions of libgevb.so (sm_100a only)
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>                 // CUtensorMap; the encoder is fetched from the driver at run time, nothing links libcuda
#include <cufft.h>
#include <nccl.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/gevb.h"

// ---------------------------------------------------------------- errors -----
void gevb_set_error(const char * fmt, ...);
// 3-D tensor map (TMA) over one component of a real field, double [nzl + 2][N][N], with the given box; returns non-zero with the error set
struct gevb_ctx;
int gevb_tensor_map_3d(gevb_ctx * c, CUtensorMap * map, const double * base, int box_x, int box_y, int box_z);

// ---------------------------------------------------------------- tuning knobs (ctx.cu)
enum { TUNE_GEODESIC_VARIANT = 0, TUNE_FFT_EXCHANGE, TUNE_FFT_OVERLAP, TUNE_FFT_DECOMPOSED, TUNE_DEPOSIT_VARIANT, TUNE_FFT_L2_PLANES, TUNE_REBIN_VARIANT, TUNE_FFT_FUSED, TUNE_PEER_COMM, TUNE_GEODESIC_TMA, TUNE_TMA_L2_PROMOTION, TUNE_FFT_XPASS, GEVB_NTUNE };
int gevb_tune(int knob);

// ---------------------------------------------------------------- NCCL (dlopen'ed, see nccl_dl.cu)
struct GevbNccl
{
	ncclResult_t (*GetUniqueId)(ncclUniqueId *);
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	const char * (*GetErrorString)(ncclResult_t);
};
GevbNccl * gevb_nccl();
int gevb_nccl_load();
#define ncclGetUniqueId gevb_nccl()->GetUniqueId
#define ncclCommInitRank gevb_nccl()->CommInitRank
#define ncclCommDestroy gevb_nccl()->CommDestroy
#define ncclSend gevb_nccl()->Send
#define ncclRecv gevb_nccl()->Recv
#define ncclAllReduce gevb_nccl()->AllReduce
#define ncclGroupStart gevb_nccl()->GroupStart
#define ncclGroupEnd gevb_nccl()->GroupEnd
#define ncclGetErrorString gevb_nccl()->GetErrorString

#define GEVB_FAIL(...) do { gevb_set_error(__VA_ARGS__); return 1; } while (0)
#define GEVB_CHECK_ARG(cond, ...) do { if (!(cond)) GEVB_FAIL(__VA_ARGS__); } while (0)
#define CUDA_TRY(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) GEVB_FAIL("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(e_), cudaGetErrorString(e_)); } while (0)
#define CUFFT_TRY(expr) do { cufftResult r_ = (expr); if (r_ != CUFFT_SUCCESS) GEVB_FAIL("%s:%d cuFFT error %d", __FILE__, __LINE__, (int) r_); } while (0)
#define NCCL_TRY(expr) do { ncclResult_t r_ = (expr); if (r_ != ncclSuccess) GEVB_FAIL("%s:%d NCCL error %s", __FILE__, __LINE__, ncclGetErrorString(r_)); } while (0)
#define GEVB_TRY(expr) do { int r_ = (expr); if (r_ != 0) return r_; } while (0)
#define KERNEL_CHECK(ctx) do { (ctx)->launches++; CUDA_TRY(cudaGetLastError()); } while (0)

// ---------------------------------------------------------------- timing (timing.cu)
enum
{
	CLS_INIT = 0, CLS_T00, CLS_TIJ, CLS_T00_TIJ, CLS_T0I, CLS_COMM, CLS_SUM, CLS_PREP_SCALAR, CLS_PREP_TENSOR, CLS_FFT_FWD, CLS_FFT_BWD,
	CLS_POISSON, CLS_FTSCALAR, CLS_EVOLVE, CLS_FTVECTOR, CLS_FTTENSOR, CLS_HALO, CLS_KICK, CLS_DRIFT, CLS_KICK_DRIFT, CLS_SORT, CLS_SPECTRUM,
	CLS_MIGRATE, CLS_FFT_A2A, CLS_FFT_TRANSPOSE, CLS_FTSCALAR_EVOLVE, GEVB_NCLS
};
struct GevbTimer
{
	bool on = false;
	std::vector<cudaEvent_t> ev;
	std::vector<int> cls;
	std::vector<size_t> open;  // indices of the begun, not yet ended pairs (scopes nest: the FFT times its exchange inside itself)
	size_t used = 0;
};
#define GEVB_MAX_RANKS 16
#define GEVB_NXEV 32
struct gevb_ctx;
struct gevb_plan;
void gevb_timer_begin(gevb_ctx * c, int cls);
void gevb_timer_end(gevb_ctx * c);
struct Timed
{
	gevb_ctx * c;
	Timed(gevb_ctx * ctx, int cls) : c(ctx) { gevb_timer_begin(c, cls); }
	~Timed() { gevb_timer_end(c); }
};

// ---------------------------------------------------------------- context ----
struct gevb_ctx
{
	GevbTimer * timer;
	int N, nh;                 // lattice points per dimension, N/2+1
	int device, rank, nranks;
	int z0, nzl;               // z-slab of real space owned by this rank
	int ky0, nkyl;             // ky-slab of Fourier space owned by this rank (nranks > 1)
	int num_sms;
	cudaStream_t stream;
	ncclComm_t comm;
	bool have_comm;
	double * d_gridk2;         // (2N sin(pi i/N))^2, host-computed exactly as gevolution.hpp:224-226
	double2 * d_kshift;        // 2N sin(pi i/N) e^{-i pi i/N}
	void * scratch;            // grow-only device scratch (sort temp, reductions, FFT staging)
	size_t scratch_bytes;
	void * scratch2;
	size_t scratch2_bytes;
	double * d_red;            // small device reduction buffer (4096 doubles)
	double * h_red;            // pinned mirror
	int64_t launches;
	// slab FFT exchange over peer memory (fft.cu): xchg[b][r] is buffer b of rank r as mapped into this process
	// (cudaIpc; r == rank is the local allocation).  Two buffers alternate so that one barrier per transform suffices.
	void * xchg[2][GEVB_MAX_RANKS];
	size_t xchg_bytes;
	uint64_t xchg_epoch;
	int xchg_state;            // 0 not tried, 1 mapped, -1 unavailable (NCCL exchange is used)
	int * d_barrier;
	cudaStream_t xstream;      // the pushes of component k run here while the local transform of component k+1 runs on `stream`
	cudaEvent_t xev[GEVB_NXEV + 1];   // [0..GEVB_NXEV) a piece's local transform is done (used round robin); [GEVB_NXEV] exchange + barrier are done
	// ghost planes, deposit folds, particle migration and rank barriers over peer memory (peer.cu): pc[r] is the
	// communication buffer of rank r as mapped into this process (cudaIpc), pf[r] its flag page
	void * pc[GEVB_MAX_RANKS];
	unsigned long long * pf[GEVB_MAX_RANKS];
	size_t pc_bytes, pc_plane_doubles, pc_mig_cap;     // whole buffer; doubles of the plane region of one slot; particles per direction of a slot
	uint64_t pc_seq, pf_epoch;                         // collective operations so far (slot = pc_seq & 1); barriers so far
	int peer_state;                                    // 0 not tried, 1 mapped, -1 unavailable (NCCL point-to-point is used)
	int * d_peer_err;                                  // set by a barrier that timed out
	size_t plane() const { return (size_t) N * N; }
	size_t real_comp_stride() const { return (size_t) (nzl + 2) * N * N; }
	size_t cplx_comp_stride() const { return nranks == 1 ? (size_t) N * N * nh : (size_t) nkyl * nh * N; }
};

int gevb_ctx_scratch(gevb_ctx * ctx, size_t bytes, void ** out);
// ---- peer.cu: communication over peer memory (one node, NVLink / NVSwitch) ---------------------------------------
int gevb_peer_setup(gevb_ctx * ctx);                       // collective; falls back silently (peer_state = -1) where peer mapping is unavailable
void gevb_peer_release(gevb_ctx * ctx);
bool gevb_peer_on(const gevb_ctx * ctx);                   // mapped and not switched off by the tuning knob peer_comm
int gevb_peer_barrier(gevb_ctx * ctx, cudaStream_t stream);   // stream-ordered barrier of all ranks: flags in peer memory, one small kernel
int gevb_peer_share(gevb_ctx * ctx, void * mine, void ** all, cudaStream_t stream);   // collective: IPC handles of `mine` to every rank, mapped into all[r]
int gevb_peer_halo(gevb_field * f);                        // Field::updateHalo across ranks
int gevb_peer_fold(gevb_field * f);                        // *_comm: upper ghost plane added into the next rank's first bulk plane
struct PeerSlot { double * base[GEVB_MAX_RANKS]; };        // the current slot of every rank's buffer
double * gevb_peer_slot(gevb_ctx * ctx, int rank, int slot);
int gevb_peer_check(gevb_ctx * ctx);                       // after a host synchronisation: did a barrier time out?
void gevb_xchg_release(gevb_ctx * ctx);      // fft.cu: unmap / free the peer exchange buffers
int gevb_ctx_scratch2(gevb_ctx * ctx, size_t bytes, void ** out);

// ---------------------------------------------------------------- fields -----
// real : double  [ncomp][nzl+2][N][N]   plane p = local z + 1; p = 0 and p = nzl+1 are ghost planes
// cplx : double2 [ncomp][N (kz)][N (ky)][nh (kx)]          when nranks == 1  ("natural")
//        double2 [ncomp][nkyl (ky)][nh (kx)][N (kz)]       when nranks  > 1  ("slab", kz fastest)
struct gevb_field
{
	gevb_ctx * ctx;
	int kind, ncomp, symmetric;
	double * data;
	size_t comp_stride;        // in elements (double or double2)
	size_t bytes;
};

struct gevb_plan
{
	gevb_ctx * ctx;
	gevb_field * real_field, * cplx_field;
	cufftHandle fwd, bwd;      // nranks == 1: 3-D D2Z / Z2D batched over components
	cufftHandle f2d, bz1d, b2d; // nranks == 1, one component: 2-D D2Z per plane, 1-D Z2Z along z (stride N*nh), 2-D Z2D per plane
	cufftHandle fwd2d, bwd2d, z1d;   // nranks > 1: per-plane 2-D transforms + 1-D along z
	cufftHandle z1d_one;             // 1-D along z for one component (component-pipelined backward transform)
	cufftHandle fwd2d_c, z1d_c;      // nranks > 1: the same for one piece of a component (chunks > 1)
	int chunks, chunks_bwd;          // pieces per component of the exchange pipeline (forward: planes, backward: rows)
	cufftHandle f2d_c, b2d_c;        // nranks == 1: the 2-D transforms of `chunk_planes` planes at a time (tuning knob fft_l2_planes)
	cufftHandle yz2d;                // nranks == 1, own x-pass (xpass.cu): y- and z-pass as one strided 2-D Z2Z over (z, y), batched along kx; 0: not created
	int chunk_planes;                // 0: not created
	bool multi;
	bool preserve;             // backward execute keeps the Fourier field intact (default)
};

// own x-pass of the forward transform with the source preparation fused into its load (xpass.cu)
bool gevb_xpass_available(const gevb_plan * p);
int gevb_xpass_forward(gevb_plan * p, int mode, const double * phi, const double * chi, double bgmodel, double coeff, double coeff2, double coeff3, double * sum_dev);

// ---------------------------------------------------------------- particles --
// Particle order: "brick-major cell order".  The local slab is cut into bricks of 16 x 8 x 4 cells (x, y, z);
// the sort key of a particle is (brick index << 9) | (cell inside the brick), so that the particles of one
// brick are contiguous (one thread block stages the brick's field tile in shared memory) and, inside the
// brick, the particles of one cell are contiguous (deposits accumulate per cell).  16 cells along x: the 16
// lanes of a half-warp that hold consecutive cells touch 16 consecutive doubles of a tile row, which is
// free of shared-memory bank conflicts for 8-byte accesses.  cell_start[key] is the exclusive prefix sum of
// the per-cell counts in key order (the counting sort's offsets), kept valid at all times.
#define GEVB_BX 16
#define GEVB_BY 8
#define GEVB_BZ 4
#define GEVB_BX_BITS 4
#define GEVB_BY_BITS 3
#define GEVB_BZ_BITS 2
#define GEVB_BRICK_CELLS 512
#define GEVB_BRICK_BITS 9
// Bricks are numbered super-brick by super-brick (4 x 4 x 8 bricks = 64 x 32 x 32 cells): the persistent kernels
// walk the bricks in index order, so the few hundred bricks in flight at any time form a compact region of the
// lattice and the tile aprons they share (deposit flushes, field tiles) meet in L2 instead of HBM.
#define GEVB_SX_BITS 2
#define GEVB_SY_BITS 2
#define GEVB_SZ_BITS 3
#define GEVB_SUPER_BITS (GEVB_SX_BITS + GEVB_SY_BITS + GEVB_SZ_BITS)
#define GEVB_INVALID_KEY 0xffffffffu
struct BrickGeom
{
	int N, nzl, z0;
	int nsx, nsy, nsz;         // super-bricks per dimension (the last ones may be partial: their missing bricks stay empty)
	int sx_shift, sy_shift;    // log2(nsx), log2(nsy) when both are powers of two (the usual lattice sizes), else -1: no integer division per brick
	uint32_t nbricks, ncells;  // nbricks = nsx nsy nsz 128, ncells = nbricks * 512
};
__host__ __device__ __forceinline__ uint32_t brick_key(const BrickGeom & G, int cx, int cy, int czl)
{
	const uint32_t bx = (uint32_t) cx >> GEVB_BX_BITS, by = (uint32_t) cy >> GEVB_BY_BITS, bz = (uint32_t) czl >> GEVB_BZ_BITS;
	const uint32_t super = ((bz >> GEVB_SZ_BITS) * G.nsy + (by >> GEVB_SY_BITS)) * G.nsx + (bx >> GEVB_SX_BITS);
	const uint32_t local = (((bz & ((1u << GEVB_SZ_BITS) - 1)) << GEVB_SY_BITS | (by & ((1u << GEVB_SY_BITS) - 1))) << GEVB_SX_BITS) | (bx & ((1u << GEVB_SX_BITS) - 1));
	const uint32_t b = (super << GEVB_SUPER_BITS) | local;
	return (b << GEVB_BRICK_BITS) | (uint32_t) (((czl & (GEVB_BZ - 1)) << (GEVB_BX_BITS + GEVB_BY_BITS)) | ((cy & (GEVB_BY - 1)) << GEVB_BX_BITS) | (cx & (GEVB_BX - 1)));
}
__host__ __device__ __forceinline__ void brick_origin(const BrickGeom & G, uint32_t brick, int & x0, int & y0, int & zl0)
{
	const uint32_t local = brick & ((1u << GEVB_SUPER_BITS) - 1), super = brick >> GEVB_SUPER_BITS;
	uint32_t sx, sy, sz;
	if (G.sx_shift >= 0) { sx = super & ((1u << G.sx_shift) - 1); sy = (super >> G.sx_shift) & ((1u << G.sy_shift) - 1); sz = super >> (G.sx_shift + G.sy_shift); }
	else { const uint32_t r = super / G.nsx; sx = super % G.nsx; sy = r % G.nsy; sz = r / G.nsy; }
	const uint32_t lx = local & ((1u << GEVB_SX_BITS) - 1), ly = (local >> GEVB_SX_BITS) & ((1u << GEVB_SY_BITS) - 1), lz = local >> (GEVB_SX_BITS + GEVB_SY_BITS);
	x0 = (int) (((sx << GEVB_SX_BITS) | lx) << GEVB_BX_BITS);
	y0 = (int) (((sy << GEVB_SY_BITS) | ly) << GEVB_BY_BITS);
	zl0 = (int) (((sz << GEVB_SZ_BITS) | lz) << GEVB_BZ_BITS);
}
// cell = floor(pos/dx) clamped into the lattice (LATfield2 filing rule, reference uses at gevolution.hpp:979-983)
__device__ __forceinline__ int cell_of(double p, double dx, int N)
{
	int c = (int) floor(p / dx);
	c = c >= N ? N - 1 : c;
	return c < 0 ? 0 : c;
}

struct gevb_pcls
{
	gevb_ctx * ctx;
	double mass;
	int64_t n, cap;
	// brick-major cell-sorted structure of arrays; [cur] is live, [1-cur] is the re-bin's output buffer
	double * x[2], * y[2], * z[2], * qx[2], * qy[2], * qz[2];
	int64_t * id[2];
	uint32_t * key;            // sort key of each live particle after a drift (input of the re-bin)
	uint32_t * rank;           // its rank among the particles of the same new cell (what the histogram's atomicAdd returned)
	uint32_t * cell_count;     // [ncells + 1] histogram of keys; all zero between re-bins
	uint32_t * cell_start;     // [ncells + 1] exclusive prefix sum; cell_start[ncells] == n
	BrickGeom geom;
	int cur;
	const unsigned long long * d_nin;   // when set, the re-bin's move reads the number of records from here (the host passes an upper bound)
};

int gevb_pcls_reserve(gevb_pcls * p, int64_t cap);
// counting sort of the first n_in live particles by p->key (entries with GEVB_INVALID_KEY are dropped);
// hist_valid: cell_count already holds the histogram of the keys (the drift kernel accumulates it)
// (with tuning knob rebin_variant = 1 the producers of the histogram also keep the value each atomicAdd returned in
// p->rank, and the move needs no atomic: slot = cell_start[key] + rank)
int gevb_pcls_rebin(gevb_pcls * p, int64_t n_in, int64_t n_out, bool hist_valid);

// Fourier-space index decode shared by all k-kernels
struct KLayout
{
	int N, nh, slab, ky0, nkyl;
	size_t sites;              // complex sites per component on this rank
	uint64_t mN, mnh;          // ceil(2^40 / N), ceil(2^40 / nh): i / d == (i * m) >> 40 exactly for i * d < 2^40 (no integer division per site)
};
static inline KLayout make_klayout(const gevb_ctx * c)
{
	KLayout L;
	L.N = c->N; L.nh = c->nh; L.slab = c->nranks > 1; L.ky0 = c->ky0; L.nkyl = c->nkyl;
	L.sites = c->cplx_comp_stride();
	L.mN = ((1ull << 40) + (uint64_t) L.N - 1) / (uint64_t) L.N;
	L.mnh = ((1ull << 40) + (uint64_t) L.nh - 1) / (uint64_t) L.nh;
	return L;
}
__device__ __forceinline__ void k_decode(const KLayout & L, size_t i, int & kx, int & ky, int & kz)
{
	// i < 2^31 (checked where the fields are created): quotients by multiply-shift, see KLayout
	const uint32_t ii = (uint32_t) i;
	if (L.slab)
	{
		const uint32_t r = (uint32_t) (((uint64_t) ii * L.mN) >> 40), q = (uint32_t) (((uint64_t) r * L.mnh) >> 40);
		kz = (int) (ii - r * (uint32_t) L.N);
		kx = (int) (r - q * (uint32_t) L.nh); ky = (int) q + L.ky0;
	}
	else
	{
		const uint32_t r = (uint32_t) (((uint64_t) ii * L.mnh) >> 40), q = (uint32_t) (((uint64_t) r * L.mN) >> 40);
		kx = (int) (ii - r * (uint32_t) L.nh);
		ky = (int) (r - q * (uint32_t) L.N); kz = (int) q;
	}
}

static inline int gevb_grid(const gevb_ctx * ctx, size_t work, int block, int per_sm = 8)
{
	size_t want = (work + block - 1) / block;
	size_t cap = (size_t) ctx->num_sms * per_sm;
	if (want < 1) want = 1;
	return (int) (want < cap ? want : cap);
}
