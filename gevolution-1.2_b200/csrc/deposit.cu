// deposit.cu -- particle -> mesh projections (CIC / staggered CIC-NGP)
//
//   projection_T00_project       gevolution.hpp:927-1022
//   projection_T0i_project       gevolution.hpp:1046-1147
//   projection_Tij_project       gevolution.hpp:1173-1297
//   scalarProjectionCIC_project  LATfield2 (main.cpp:402), plain CIC
//
// Persistent thread blocks walk the bricks of 16 x 8 x 4 cells (particles are
// stored brick by brick and, inside a brick, cell by cell -- gevb_internal.cuh).
// A block owns a 17 x 9 x 5-site shared-memory tile per target component (the
// brick's sites plus the upper apron the CIC cloud reaches); the brick's slice of
// cell_start[] and its phi tile arrive by asynchronous copies (LDGSTS) one brick
// ahead, the particles by register prefetch one batch ahead.
//
//   1. one thread per particle (coalesced 48-byte SoA loads); the per-particle
//      factors (weights, e, f, m q_i q_j / e) stay in registers.
//   2. the tile is updated in eight corner phases: in one phase every particle adds
//      to the same corner of its own cell, so particles of different cells touch
//      different sites and the shared-memory update is a plain read-add-write.
//      Particles that share a cell are neighbours in the sorted order: their
//      contributions are combined by a segmented warp reduction (shuffles, only in
//      warps that hold such particles) and the first lane of the segment writes --
//      the GPU form of the reference's per-cell "localCube" accumulators
//      (gevolution.hpp:953,1199-1200).  Only a cell whose particles straddle two
//      warps needs a shared-memory atomic.  Work per thread does not depend on
//      the clustering state.
//   3. the tile is flushed to the field in HBM with FP64 reductions: one bulk
//      reduction of the TMA unit per 16-site row and component (UBLKRED, 315 per
//      brick) plus one RED.ADD.F64 for the apron site of the row -- instead of 38
//      reductions per particle; bricks whose rows wrap in x flush site by site.
//
// x and y wrap by index arithmetic at flush time; z+1 of the last local plane
// lands in the upper ghost plane which gevb_projection_comm folds into the next rank.
#include "gevb_internal.cuh"

namespace {

#define DX (GEVB_BX + 1)
#define DY (GEVB_BY + 1)
#define DZ (GEVB_BZ + 1)
#define DT_SITES (DX * DY * DZ)
#define DEP_THREADS 256

enum { DEP_T00 = 0, DEP_TIJ = 1, DEP_T00_TIJ = 2, DEP_T0I = 3 };
// flavours of k_deposit (tuning knob deposit_variant: 4 = DEPF_BULK, the default; 0 = none, the round-1 kernel):
//   DEPF_BULK   the tile is flushed row by row with bulk reductions of the TMA unit (cp.reduce.async.bulk ... add.f64, UBLKRED: one
//               per row of 16 sites and component, 315 per brick) instead of one RED per site (5355 per brick); rows of the
//               accumulator tile are padded to 18 sites so that each starts on a 16-byte boundary.
// Two more were measured and dropped (profiles/round2_v08): a split phase barrier (mbarrier arrive after the updates of phase k,
// wait before those of phase k + 1: +7 %) and turn-taking of a cell's lanes instead of the shuffle sum (2 x slower).
//   DEPF_LOOP   (deposit_variant 6, with DEPF_BULK) the shuffle steps of the segmented sum are a run-time loop: one copy of the
//               eight phases in the instruction stream instead of four (0, 1, 2, 5 steps)
//   DEPF_SPILL  (deposit_variant 8 with DEPF_BULK, 10 with both) no shared-memory atomics inside the phases.  The only lanes that could
//               meet another warp on a site are those whose cell began in the previous warp of the batch; they accumulate into a
//               spare "cell" of their warp in front of the tile of every component (acc_base) with the same plain
//               read-add-write as everyone else, and one extra step per batch merges the spare cells into the real ones
//   DEPF_TMA    (deposit_variant 18, with the three above) the whole tile of a component is flushed by ONE tensor reduction of the TMA unit
//               (cp.reduce.async.bulk.tensor.3d ... add, UTMAREDG) through a 3-D tensor map over the field [nzl+2][N][N]: 7 per brick
//               instead of 315 row reductions + 315 apron REDs.  Sites past the lattice edge are dropped by the unit (the planes above a
//               partial brick, by construction never to be flushed -- and the apron of a brick at the upper x or y edge, which wraps
//               around and is flushed with REDs); the padding column and the (zero) spare rows add 0 to the neighbouring brick's sites
enum { DEPF_BULK = 1, DEPF_LOOP = 2, DEPF_SPILL = 4, DEPF_TMA = 8 };
struct DMaps { CUtensorMap m[7]; };                      // one tensor map per target component
__host__ __device__ constexpr int acc_ax(int flags) { return (flags & DEPF_BULK) ? DX + 1 : DX; }
__host__ __device__ constexpr int acc_ay(int flags) { return DY; }                                              // rows of a plane of the accumulator tile
// DEPF_SPILL: every component is preceded by 208 doubles that hold the warps' spare cells.  The spare cell of warp w starts 2 w doubles
// into them and uses the strides of the tile (1, 18, 162), so its corners lie at 2w + {0, 1, 18, 19, 162, 163, 180, 181}: inside the 208,
// disjoint between warps, and outside the dense 18 x 9 x 5 box that the TMA unit reads (which starts 1664 bytes = 13 x 128 in)
__host__ __device__ constexpr int acc_base(int flags) { return (flags & DEPF_SPILL) ? 208 : 0; }
__host__ __device__ constexpr int acc_sites(int flags)                                                          // component stride
{
	return (flags & (DEPF_TMA | DEPF_SPILL)) ? (acc_base(flags) + acc_ax(flags) * acc_ay(flags) * DZ + 15) / 16 * 16 : acc_ax(flags) * acc_ay(flags) * DZ;
}
__host__ __device__ constexpr int acc_corner(int flags, int k) { return ((k >> 2) & 1) + ((k >> 1) & 1) * acc_ax(flags) + (k & 1) * acc_ax(flags) * acc_ay(flags); }

// tile components per projection; Tij components follow the field order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
__host__ __device__ constexpr int dep_ncomp(int what) { return what == DEP_T00 ? 1 : what == DEP_TIJ ? 6 : what == DEP_T00_TIJ ? 7 : 3; }
// corner index 4X + 2Y + Z (gevolution.hpp:953) -> offset inside the tile
__host__ __device__ constexpr int corner_offset(int k) { return ((k >> 2) & 1) + ((k >> 1) & 1) * DX + (k & 1) * DX * DY; }

struct DParams
{
	BrickGeom G;
	int pow2;
	double dx, rN, a, mass;
	const uint32_t * cell_start;
	const double * x, * y, * z, * qx, * qy, * qz;
	const double * phi;
	double * out[7];           // target component pointers in tile-component order
};

__device__ __forceinline__ void cp_async8(void * smem_dst, const void * gmem_src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(void * smem_dst, const void * gmem_src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ uint32_t smem_addr(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long * bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long * bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long * bar, uint32_t parity)
{
	uint32_t done = 0;
	while (!done)
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void * smem_dst, const void * gmem_src, uint32_t bytes, unsigned long long * bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
		:: "r"(smem_addr(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// bulk reduction (TMA unit, UBLKRED): global[0 .. bytes) += shared[0 .. bytes) as FP64 adds; 16-byte aligned on both sides
__device__ __forceinline__ void bulk_add_f64(double * gmem_dst, const double * smem_src, uint32_t bytes)
{
	asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" :: "l"(gmem_dst), "r"(smem_addr(smem_src)), "r"(bytes) : "memory");
}
// tensor reduction (TMA unit): the box of the tensor map at (x, y, z) += the dense box at smem_src (128-byte aligned)
__device__ __forceinline__ void tensor_add_3d(const CUtensorMap * map, const double * smem_src, int x, int y, int z)
{
	asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
		:: "l"((unsigned long long) map), "r"(x), "r"(y), "r"(z), "r"(smem_addr(smem_src)) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// periodic wrap of an index in [0, 2N) (a tile reaches one site past the brick; bricks of a partial super-brick lie past N)
__device__ __forceinline__ int wrap_up(int v, int N) { v = v >= N ? v - N : v; return v >= N ? v % N : v; }

#define DEP_CELLTAB (GEVB_BRICK_CELLS + 4)                      // 513 prefix sums of a brick, padded
#define DEP_STAGE_DOUBLES (DT_SITES + DEP_CELLTAB / 2)          // one pipeline stage: phi tile, cell table

// asynchronous copy (LDGSTS) of a brick's slice of cell_start[] and its phi tile.  For the tile a thread keeps its
// (tx, ty) column and walks z.
template <bool HAS_PHI>
__device__ __forceinline__ void stage_brick(const DParams & D, uint32_t brick, uint32_t first, uint32_t last, double * stage)
{
	const BrickGeom & G = D.G;
	double * tphi = stage;
	uint32_t * ctab = (uint32_t *) (stage + DT_SITES);
	if (first == last) return;
	const uint32_t * src = D.cell_start + (size_t) brick * GEVB_BRICK_CELLS;
	for (int k = threadIdx.x; k <= GEVB_BRICK_CELLS; k += DEP_THREADS) cp_async4(ctab + k, src + k);
	if (HAS_PHI && threadIdx.x < DX * DY)
	{
		int x0, y0, zl0;
		brick_origin(G, brick, x0, y0, zl0);
		const int tx = threadIdx.x % DX, ty = threadIdx.x / DX;
		const size_t gcol = (size_t) wrap_up(y0 + ty, G.N) * G.N + wrap_up(x0 + tx, G.N);
		#pragma unroll
		for (int tz = 0; tz < DZ; tz++)
		{
			const int plane = zl0 + tz + 1;
			if (plane > G.nzl + 1) break;                       // partial brick at the top of the slab: never read
			cp_async8(tphi + (tz * DY + ty) * DX + tx, D.phi + (size_t) plane * G.N * G.N + gcol);
		}
	}
}

__device__ __forceinline__ void brick_range(const DParams & D, uint32_t b, uint32_t & first, uint32_t & last)
{
	first = last = 0;
	if (b < D.G.nbricks) { first = __ldg(D.cell_start + (size_t) b * GEVB_BRICK_CELLS); last = __ldg(D.cell_start + (size_t) (b + 1) * GEVB_BRICK_CELLS); }
}

// per-particle factors, computed once and kept in registers through the eight phases
struct PInv
{
	double wx[2], wy[2], wz[2];    // CIC weights: [0] = 1 - up ("down"), [1] = up     gevolution.hpp:979-983
	double eT, fT;                 // T00: mass * (e + f phi)                           :989-1009
	double f;                      // Tij: 4 + a^2 / (q^2 + a^2)                        :1238
	double wd[3], wo[3];           // Tij: mass q_d^2 / e ; mass q_i q_j / e (01, 02, 12)  :1243,:1262-1270
	double m[3];                   // T0i: mass q_i                                     :1107-1121
};

template <int WHAT, bool HAS_PHI>
__device__ __forceinline__ void particle_factors(PInv & I, const DParams & D, const double * pv, int cx, int cy, int cz)
{
	double up[3];
	const double refx = cx * D.dx, refy = cy * D.dx, refz = cz * D.dx;                                     // referPos, :963
	if (D.pow2) { up[0] = (pv[0] - refx) * D.rN; up[1] = (pv[1] - refy) * D.rN; up[2] = (pv[2] - refz) * D.rN; }   // == / dx exactly (dx = 2^-k)
	else { up[0] = (pv[0] - refx) / D.dx; up[1] = (pv[1] - refy) / D.dx; up[2] = (pv[2] - refz) / D.dx; }          // :981 / :1101 / :1231
	I.wx[1] = up[0]; I.wy[1] = up[1]; I.wz[1] = up[2];
	I.wx[0] = 1.0 - up[0]; I.wy[0] = 1.0 - up[1]; I.wz[0] = 1.0 - up[2];                                     // :982
	const double q0 = pv[3], q1 = pv[4], q2 = pv[5];
	if (WHAT == DEP_T0I) { I.m[0] = D.mass * q0; I.m[1] = D.mass * q1; I.m[2] = D.mass * q2; return; }      // :1107,:1114,:1121
	const double qsq = q0 * q0 + q1 * q1 + q2 * q2;
	// neither a square root nor a division per particle: r = (q^2 + a^2)^(-1/2) gives 1/e = r and e = (q^2 + a^2) r; the
	// reference's quotients x / e become x * r (a few ulp, 1e-16 relative, against a 1e-10 tolerance)
	const double inv_e = rsqrt(qsq + D.a * D.a);
	const double e = (qsq + D.a * D.a) * inv_e;                                    // :990 / :1237
	if (WHAT == DEP_T00 || WHAT == DEP_T00_TIJ)
	{
		I.eT = D.a * D.mass; I.fT = 0.;
		if (HAS_PHI) { I.eT = e * D.mass; I.fT = (3. * e + qsq * inv_e) * D.mass; }   // :989-991; mass applied at write-back in the reference (:1012)
	}
	if (WHAT == DEP_TIJ || WHAT == DEP_T00_TIJ)
	{
		I.f = 4. + D.a * D.a * inv_e * inv_e;                                      // :1238
		const double me = D.mass * inv_e;
		I.wd[0] = me * q0 * q0; I.wd[1] = me * q1 * q1; I.wd[2] = me * q2 * q2;    // :1243
		I.wo[0] = me * q0 * q1; I.wo[1] = me * q0 * q2; I.wo[2] = me * q1 * q2;    // :1262,:1266,:1270
	}
}

// how a lane writes its contributions
struct Writer
{
	double * tile;                 // tile + site of the particle's cell
	int after, steps;              // lanes after this one in the same cell (within the warp); shuffle steps of the segmented sum
	bool write;                    // this lane writes: alone in its cell, or first lane of its cell's segment in this warp
	bool plain;                    // no other warp holds particles of this cell: plain read-add-write is safe
};

// tile component of the j-th contribution of corner K (j = -1: how many contributions corner K has)
__host__ __device__ constexpr int phase_comp(int what, int K, int j)
{
	const int X = (K >> 2) & 1, Y = (K >> 1) & 1, Z = K & 1;
	int comps[7] = {0, 0, 0, 0, 0, 0, 0};
	int n = 0;
	if (what == DEP_T0I)
	{
		if (X == 0) comps[n++] = 0;
		if (Y == 0) comps[n++] = 1;
		if (Z == 0) comps[n++] = 2;
	}
	else
	{
		const int o = what == DEP_T00_TIJ ? 1 : 0;
		if (what != DEP_TIJ) comps[n++] = 0;
		if (what != DEP_T00)
		{
			comps[n++] = o; comps[n++] = o + 3; comps[n++] = o + 5;
			if (X == 0 && Y == 0) comps[n++] = o + 1;
			if (X == 0 && Z == 0) comps[n++] = o + 2;
			if (Y == 0 && Z == 0) comps[n++] = o + 4;
		}
	}
	return j < 0 ? n : comps[j];
}

// contributions of one particle to corner K of its cell, in the order of phase_comp()
// ph: the phi tile at the particle's cell (corner k at ph[corner_offset(k)]), or, with CORNERS, the eight corner values themselves
template <int WHAT, bool HAS_PHI, int K, bool CORNERS = false>
__device__ __forceinline__ void phase_values(const PInv & I, const double * ph, double * v)
{
	constexpr int X = (K >> 2) & 1, Y = (K >> 1) & 1, Z = K & 1;
	#define PHI(k) (HAS_PHI ? ph[CORNERS ? (k) : corner_offset(k)] : 0.)
	int n = 0;
	if (WHAT == DEP_T0I)
	{
		// component i is NGP along i and CIC across (:1109-1126); edge factor 1 + phi(x) + phi(x + e_i) (:1129-1144)
		if (X == 0) v[n++] = I.m[0] * I.wy[Y] * I.wz[Z] * (1. + PHI(K) + PHI(K + 4));
		if (Y == 0) v[n++] = I.m[1] * I.wx[X] * I.wz[Z] * (1. + PHI(K) + PHI(K + 2));
		if (Z == 0) v[n++] = I.m[2] * I.wx[X] * I.wy[Y] * (1. + PHI(K) + PHI(K + 1));
		return;
	}
	const double w = I.wx[X] * I.wy[Y] * I.wz[Z];
	const double p = PHI(K);
	if (WHAT == DEP_T00 || WHAT == DEP_T00_TIJ) v[n++] = w * (I.eT + I.fT * p);                             // :995-1019
	if (WHAT == DEP_TIJ || WHAT == DEP_T00_TIJ)
	{
		const double g = w * (1. + I.f * p);
		v[n++] = I.wd[0] * g; v[n++] = I.wd[1] * g; v[n++] = I.wd[2] * g;                                   // :1245-1259
		// off-diagonal (i,j) lives on the plaquette centre: CIC along the third axis only, phi averaged over the plaquette
		if (X == 0 && Y == 0) v[n++] = I.wo[0] * I.wz[Z] * (1. + I.f * 0.25 * (PHI(Z) + PHI(2 + Z) + PHI(4 + Z) + PHI(6 + Z)));                       // :1263-1264, at x (Z=0) and x+e2 (Z=1) :1279,:1290
		if (X == 0 && Z == 0) v[n++] = I.wo[1] * I.wy[Y] * (1. + I.f * 0.25 * (PHI(2 * Y) + PHI(2 * Y + 1) + PHI(2 * Y + 4) + PHI(2 * Y + 5)));       // :1267-1268, x and x+e1
		if (Y == 0 && Z == 0) v[n++] = I.wo[2] * I.wx[X] * (1. + I.f * 0.25 * (PHI(4 * X) + PHI(4 * X + 1) + PHI(4 * X + 2) + PHI(4 * X + 3)));       // :1271-1272, x and x+e0
	}
	#undef PHI
}

// one corner phase of a batch.  Particles of different cells update different sites, so the update is a plain
// read-add-write.  Particles of one cell are contiguous lanes: their contributions are summed by a segmented
// shuffle reduction and the first lane of the segment writes -- with a shared-memory atomic if the cell's
// particles straddle warps (the only case in which two warps can meet on a site within a phase).
template <int WHAT, int K, int FLAGS>
__device__ __forceinline__ void tile_add(const Writer & W, const double * v)
{
	constexpr int NV = phase_comp(WHAT, K, -1);
	constexpr int ASITES = acc_sites(FLAGS);
	double * p = W.tile + acc_corner(FLAGS, K);
	if ((FLAGS & DEPF_SPILL) || W.plain)
	{
		#pragma unroll
		for (int j = 0; j < NV; j++) p[phase_comp(WHAT, K, j) * ASITES] += v[j];
	}
	else
	{
		#pragma unroll
		for (int j = 0; j < NV; j++) atomicAdd(p + phase_comp(WHAT, K, j) * ASITES, v[j]);
	}
}

template <int WHAT, bool HAS_PHI, int K, int STEPS, int FLAGS>
__device__ __forceinline__ int phase(const PInv & I, const Writer & W, const double * ph, int vote, bool solo)
{
	constexpr int NV = phase_comp(WHAT, K, -1);
	double v[NV > 0 ? NV : 1];
	phase_values<WHAT, HAS_PHI, K>(I, ph, v);
	if (STEPS >= 0)
	{
		// v += take * t with take = 1 or 0 is the same sum as a conditional add (exact products), one DFMA instead of two selects and a DADD
		#pragma unroll
		for (int s = 0; s < STEPS; s++)
		{
			const double take = (1 << s) <= W.after ? 1. : 0.;
			#pragma unroll
			for (int j = 0; j < NV; j++) v[j] = fma(__shfl_down_sync(0xffffffffu, v[j], 1 << s), take, v[j]);
		}
	}
	else
	{
		#pragma unroll 1
		for (int s = 0; s < W.steps; s++)
		{
			const double take = (1 << s) <= W.after ? 1. : 0.;
			#pragma unroll
			for (int j = 0; j < NV; j++) v[j] = fma(__shfl_down_sync(0xffffffffu, v[j], 1 << s), take, v[j]);
		}
	}
	if (W.write) tile_add<WHAT, K, FLAGS>(W, v);
	if ((FLAGS & DEPF_LOOP) && solo) { __syncwarp(); return 0; }            // the only warp at work: its own order is enough
	if ((FLAGS & DEPF_SPILL) && K == 7) return __syncthreads_or(vote);      // the barrier also tells whether any warp used its spare cell
	__syncthreads();
	return 0;
}

// cell = floor(pos/dx) clamped into the lattice; pos * N is bit-identical to pos / dx for power-of-two N
__device__ __forceinline__ int cell_scaled(const DParams & D, double p)
{
	int c = (int) floor(D.pow2 ? p * D.rN : p / D.dx);
	c = c >= D.G.N ? D.G.N - 1 : c;
	return c < 0 ? 0 : c;
}

__device__ __forceinline__ void load_particle(const DParams & D, uint32_t i, double * pv)
{
	pv[0] = D.x[i]; pv[1] = D.y[i]; pv[2] = D.z[i]; pv[3] = D.qx[i]; pv[4] = D.qy[i]; pv[5] = D.qz[i];
}

// the eight corner phases of one batch
template <int WHAT, bool HAS_PHI, int STEPS, int FLAGS>
__device__ __forceinline__ int phases(const PInv & I, const Writer & W, const double * ph, int vote, bool solo = false)
{
	phase<WHAT, HAS_PHI, 0, STEPS, FLAGS>(I, W, ph, 0, solo);
	phase<WHAT, HAS_PHI, 1, STEPS, FLAGS>(I, W, ph, 0, solo);
	phase<WHAT, HAS_PHI, 2, STEPS, FLAGS>(I, W, ph, 0, solo);
	phase<WHAT, HAS_PHI, 3, STEPS, FLAGS>(I, W, ph, 0, solo);
	phase<WHAT, HAS_PHI, 4, STEPS, FLAGS>(I, W, ph, 0, solo);
	phase<WHAT, HAS_PHI, 5, STEPS, FLAGS>(I, W, ph, 0, solo);
	phase<WHAT, HAS_PHI, 6, STEPS, FLAGS>(I, W, ph, 0, solo);
	return phase<WHAT, HAS_PHI, 7, STEPS, FLAGS>(I, W, ph, vote, solo);
}

// Persistent blocks walk the bricks with stride gridDim.x in a software pipeline: while brick k is accumulated and
// flushed, the cell table and phi tile of brick k+1 are in flight into the other shared-memory stage, the particle
// range of brick k+2 is being fetched, and the next batch of particles is already requested.
template <int WHAT, bool HAS_PHI, int FLAGS>
__global__ void __launch_bounds__(DEP_THREADS, 3) k_deposit(DParams D, const __grid_constant__ DMaps maps)
{
	constexpr int NCOMP = dep_ncomp(WHAT);
	constexpr int AX = acc_ax(FLAGS), AY = acc_ay(FLAGS), ASITES = acc_sites(FLAGS);
	extern __shared__ __align__(128) double site_smem[];
	double * smem = site_smem;
	double * tile = smem + acc_base(FLAGS);                 // [NCOMP][ASITES] accumulators (behind the spare cells of component 0)
	double * stages = smem + NCOMP * ASITES;                // [2][DEP_STAGE_DOUBLES]

	const BrickGeom & G = D.G;
	const int tcol = threadIdx.x, ttx = tcol % DX, tty = tcol / DX;      // this thread's column of the tile (flush)
	const int lane = threadIdx.x & 31;
	for (int idx = threadIdx.x; idx < NCOMP * ASITES + 2 * DEP_STAGE_DOUBLES; idx += DEP_THREADS) smem[idx] = 0.;
	__syncthreads();
	uint32_t brick = blockIdx.x;
	uint32_t first, last, nfirst, nlast, nnfirst, nnlast;
	brick_range(D, brick, first, last);
	brick_range(D, brick + gridDim.x, nfirst, nlast);
	stage_brick<HAS_PHI>(D, brick, first, last, stages);
	cp_async_commit();
	double pv[6] = {0., 0., 0., 0., 0., 0.};
	if (first + threadIdx.x < last) load_particle(D, first + threadIdx.x, pv);
	int cur = 0;
	bool pending = false;                                   // DEPF_TMA: the unit is still reading the tile of the previous brick
	while (brick < G.nbricks)
	{
		const uint32_t nbrick = brick + gridDim.x;
		brick_range(D, nbrick + gridDim.x, nnfirst, nnlast);                 // consumed at the end of this iteration
		if (nbrick < G.nbricks) stage_brick<HAS_PHI>(D, nbrick, nfirst, nlast, stages + (cur ^ 1) * DEP_STAGE_DOUBLES);
		cp_async_commit();
		if (first != last)
		{
			int x0, y0, zl0;
			brick_origin(G, brick, x0, y0, zl0);
			cp_async_wait<1>();                             // everything but the newest group: this brick's stage has landed
			__syncthreads();
			const double * tphi = stages + cur * DEP_STAGE_DOUBLES;
			const uint32_t * ctab = (const uint32_t *) (tphi + DT_SITES);

			for (uint32_t base = first; base < last; base += DEP_THREADS)
			{
				const uint32_t i = base + threadIdx.x;
				const bool valid = i < last;
				PInv I;
				Writer W;
				const double * ph;
				bool spill = false;
				int spill_to = -1;                                  // warp-uniform: site of the cell the warp's spare cell stands for
				{
					// the particle's cell (it lies in this brick: the storage order is maintained by the re-bin)
					int cx = 0, cy = 0, cz = 0, site = 0, asite = 0;
					uint32_t cfirst = i, clast = i + 1;
					if (valid)
					{
						cx = cell_scaled(D, pv[0]); cy = cell_scaled(D, pv[1]); cz = cell_scaled(D, pv[2]);
						const int sx = cx - x0, sy = cy - y0, sz = cz - G.z0 - zl0;
						site = (sz * DY + sy) * DX + sx;
						asite = (sz * AY + sy) * AX + sx;
						const int c = (sz << (GEVB_BX_BITS + GEVB_BY_BITS)) | (sy << GEVB_BX_BITS) | sx;
						cfirst = ctab[c]; clast = ctab[c + 1];
					}
					particle_factors<WHAT, HAS_PHI>(I, D, pv, cx, cy, cz);
					// lanes of this warp that share the cell form a contiguous segment [seg_lo, seg_hi)
					const uint32_t warp_lo = i - lane, warp_hi = warp_lo + 32;
					const uint32_t seg_lo = cfirst > warp_lo ? cfirst : warp_lo, seg_hi = clast < warp_hi ? clast : warp_hi;
					const int maxlen = __reduce_max_sync(0xffffffffu, (int) (seg_hi - seg_lo));
					W.tile = tile + asite; ph = tphi + site;
					if (FLAGS & DEPF_SPILL)
					{
						// the cell began in the previous warp of this batch: accumulate in the warp's spare cell
						spill = valid && cfirst < warp_lo && (threadIdx.x >> 5) != 0;
						if (spill) W.tile = smem + 2 * (threadIdx.x >> 5);
						spill_to = __shfl_sync(0xffffffffu, spill ? asite : -1, 0);
					}
					W.after = (int) (seg_hi - 1 - i);
					W.steps = maxlen > 1 ? 32 - __clz(maxlen - 1) : 0;
					W.write = valid && i == seg_lo;
					W.plain = cfirst >= warp_lo && clast <= warp_hi;
				}
				// request the next batch (of this brick, else the first batch of the next brick) while the phases run
				{
					const uint32_t nb = base + DEP_THREADS;
					const uint32_t j = nb < last ? nb + threadIdx.x : nfirst + threadIdx.x;
					if (j < (nb < last ? last : nlast)) load_particle(D, j, pv);
				}
				// shuffle steps of the segmented sum are a warp-uniform property of the batch: 0 when no two lanes share a cell;
				// a warp without particles (tail of the brick) only keeps the barriers company
				if ((FLAGS & DEPF_TMA) && pending)
				{
					// the read-out of the previous brick's tile ran behind the set-up of this batch; clear the tile before the first update
					if (threadIdx.x < NCOMP) bulk_wait_read();
					__syncthreads();
					double2 * t2 = (double2 *) smem;
					for (int idx = threadIdx.x; idx < NCOMP * ASITES / 2; idx += DEP_THREADS) t2[idx] = make_double2(0., 0.);
					__syncthreads();
					pending = false;
				}
				int spilled;
				if ((FLAGS & DEPF_LOOP) && last - base <= 32)
				{
					// the last few particles of a brick (Poisson: half of the bricks hold a little more than two batches): one warp, no barriers
					if (threadIdx.x < 32) phases<WHAT, HAS_PHI, -1, FLAGS>(I, W, ph, 0, true);
					__syncthreads();
					spilled = 0;
				}
				else if (FLAGS & DEPF_LOOP) spilled = phases<WHAT, HAS_PHI, -1, FLAGS>(I, W, ph, spill_to >= 0);
				else if (W.steps == 0) spilled = phases<WHAT, HAS_PHI, 0, FLAGS>(I, W, ph, spill_to >= 0);
				else if (W.steps == 1) spilled = phases<WHAT, HAS_PHI, 1, FLAGS>(I, W, ph, spill_to >= 0);
				else if (W.steps == 2) spilled = phases<WHAT, HAS_PHI, 2, FLAGS>(I, W, ph, spill_to >= 0);
				else spilled = phases<WHAT, HAS_PHI, 5, FLAGS>(I, W, ph, spill_to >= 0);
				if ((FLAGS & DEPF_SPILL) && spilled)
				{
					// merge the spare cells into the cells they stand for (two warps can meet on a site here: atomics, one pass of the warp)
					if (spill_to >= 0)
					{
						double * real = tile + spill_to;
						double * spare = smem + 2 * (threadIdx.x >> 5);
						for (int e = lane; e < 8 * NCOMP; e += 32)
						{
							const int k = e / NCOMP, comp = e - k * NCOMP;
							const int off = ((k >> 2) & 1) + ((k >> 1) & 1) * AX + (k & 1) * AX * AY + comp * ASITES;
							const double v = spare[off];
							if (v != 0.) { atomicAdd(real + off, v); spare[off] = 0.; }
						}
					}
					__syncthreads();
				}
			}

			// ---- flush the tile into the field in HBM; it is left zeroed for the next brick (stage `cur` is free from here on)
			if ((FLAGS & DEPF_TMA) && x0 + GEVB_BX <= G.N && y0 + GEVB_BY <= G.N)
			{
				fence_async_proxy();
				__syncthreads();
				if (threadIdx.x < NCOMP)
				{
					tensor_add_3d(&maps.m[threadIdx.x], tile + threadIdx.x * ASITES, x0, y0, zl0 + 1);
					bulk_commit();
				}
				// the apron of a brick at the upper edge of the lattice wraps around: the unit drops it, REDs take it
				const bool wrapx = x0 + GEVB_BX == G.N, wrapy = y0 + GEVB_BY == G.N;
				if (wrapx)
					for (int r = threadIdx.x; r < NCOMP * DZ * DY; r += DEP_THREADS)
					{
						const int k = r / (DZ * DY), rem = r - k * (DZ * DY), tz = rem / DY, ty = rem - tz * DY;
						const int plane = zl0 + tz + 1;
						const double v = tile[k * ASITES + (tz * AY + ty) * AX + GEVB_BX];
						if (plane <= G.nzl + 1 && v != 0.) atomicAdd(D.out[k] + (size_t) plane * G.N * G.N + (size_t) wrap_up(y0 + ty, G.N) * G.N, v);
					}
				if (wrapy)
				{
					const int nx = wrapx ? GEVB_BX : DX;
					for (int r = threadIdx.x; r < NCOMP * DZ * nx; r += DEP_THREADS)
					{
						const int k = r / (DZ * nx), rem = r - k * (DZ * nx), tz = rem / nx, tx = rem - tz * nx;
						const int plane = zl0 + tz + 1;
						const double v = tile[k * ASITES + (tz * AY + GEVB_BY) * AX + tx];
						if (plane <= G.nzl + 1 && v != 0.) atomicAdd(D.out[k] + (size_t) plane * G.N * G.N + x0 + tx, v);
					}
				}
				pending = true;                                     // completed before the first update of the next brick
			}
			else if ((FLAGS & DEPF_BULK) && x0 + GEVB_BX <= G.N)
			{
				// one bulk reduction per row of 16 sites and component (the row does not wrap in x), one RED for the apron site
				fence_async_proxy();
				__syncthreads();
				for (int r = threadIdx.x; r < NCOMP * DZ * DY; r += DEP_THREADS)
				{
					const int k = r / (DZ * DY), rem = r - k * (DZ * DY), tz = rem / DY, ty = rem - tz * DY;
					const int plane = zl0 + tz + 1;
					if (plane > G.nzl + 1) continue;                                   // partial brick at the top of the slab
					const double * row = tile + k * ASITES + (tz * AY + ty) * AX;
					double * grow = D.out[k] + (size_t) plane * G.N * G.N + (size_t) wrap_up(y0 + ty, G.N) * G.N;
					bulk_add_f64(grow + x0, row, GEVB_BX * sizeof(double));
					const double v = row[GEVB_BX];
					if (v != 0.) atomicAdd(grow + wrap_up(x0 + GEVB_BX, G.N), v);
				}
				bulk_commit();
				bulk_wait_read();                                   // the rows have been read: the tile may be cleared
				__syncthreads();
				double2 * t2 = (double2 *) smem;
				for (int idx = threadIdx.x; idx < NCOMP * ASITES / 2; idx += DEP_THREADS) t2[idx] = make_double2(0., 0.);
			}
			else
			{
				// FP64 reductions, consecutive lanes on consecutive sites of a row
				if (tcol < DX * DY)
				{
					const size_t gcol = (size_t) wrap_up(y0 + tty, G.N) * G.N + wrap_up(x0 + ttx, G.N);
					#pragma unroll
					for (int tz = 0; tz < DZ; tz++)
					{
						const int s = (tz * AY + tty) * AX + ttx;
						const size_t off = (size_t) (zl0 + tz + 1) * G.N * G.N + gcol;
						#pragma unroll
						for (int k = 0; k < NCOMP; k++)
						{
							const double v = tile[k * ASITES + s];
							if (v != 0.) { atomicAdd(D.out[k] + off, v); tile[k * ASITES + s] = 0.; }
						}
					}
				}
			}
		}
		else if (nbrick < G.nbricks && nfirst + threadIdx.x < nlast) load_particle(D, nfirst + threadIdx.x, pv);   // empty brick: nothing was prefetched
		brick = nbrick; cur ^= 1;
		first = nfirst; last = nlast; nfirst = nnfirst; nlast = nnlast;
	}
	if ((FLAGS & DEPF_TMA) && pending && threadIdx.x < NCOMP) bulk_wait_read();      // the tile must outlive the read-out
}

// =====================================================================================================================
// k_deposit_cells (tuning knob deposit_variant = 1, the default): the same projections without corner phases.
//
// The old kernel above updates a tile of lattice SITES, so two particles of neighbouring cells can meet on a site and
// every batch needs eight barrier-separated corner phases plus shared-memory CAS loops for cells split between warps
// (profiles/r1k: 4.7 barrier stalls per issue, 704 ATOMS.CAST.SPIN sites).  Here a block accumulates per CELL: every cell
// of the unit (half a brick: 16 x 8 x 2 cells) owns NACC private accumulators in shared memory, one per (corner,
// component) pair it deposits to -- the reference's localCube / localEdge arrays (gevolution.hpp:953,1075,1199-1200)
// kept for all cells of the unit at once.  Particles of different cells never touch the same word, so the accumulation
// needs no barrier and no atomic:
//
//   accumulate  each warp owns a contiguous slice of the unit's (cell-sorted) particles, one particle per lane per step.
//               Lanes that share a cell are neighbours: a segmented shuffle sum (only in warps that hold such lanes)
//               leaves the cell's contribution in the first lane of the segment, which adds it to the cell's
//               accumulators with a plain read-add-write.  A cell cut by the boundary between two warps' slices is
//               accumulated in a private column of the warp (one "head" and one "tail" column per warp) instead.
//   gather      one thread per tile site sums the up to eight cells around it, clears what it read, and issues one FP64
//               reduction (RED.ADD.F64) per component into the field in HBM; the at most 16 head / tail columns are
//               reduced into HBM directly, value by value.
//
// Two barriers per unit (about 256 particles at one particle per cell) instead of eight per batch of 256 particles, and
// they separate phases of the whole unit, not corner phases of a batch.
#define UZ 2                                           // z-layers of a unit
#define UCELLS (GEVB_BX * GEVB_BY * UZ)                // 256 cells
#define UT_SITES (DX * DY * (UZ + 1))                  // 459 tile sites (upper apron included)
#define DEP_WARPS (DEP_THREADS / 32)
#define ACC_COLS (UCELLS + 2 * DEP_WARPS)              // one column per cell + head / tail column per warp; 272 = 17 x 16: column c lives in bank pair c % 16
#define UCELLTAB (UCELLS + 4)                          // 257 prefix sums, padded
#define USTAGE_DOUBLES (UT_SITES + 1 + UCELLTAB / 2)   // one pipeline stage: phi tile, cell table

__host__ __device__ constexpr int dep_nacc(int what) { int n = 0; for (int K = 0; K < 8; K++) n += phase_comp(what, K, -1); return n; }
// accumulator of the j-th contribution of corner K: they are numbered corner by corner
__host__ __device__ constexpr int acc_index(int what, int K, int j) { int n = 0; for (int k = 0; k < K; k++) n += phase_comp(what, k, -1); return n + j; }

struct CWriter
{
	double * col;                  // acc + column of the particle's cell (or of the warp's head / tail column)
	int after;                     // lanes after this one in the same cell (within the warp step)
	bool write;                    // first lane of its cell's segment
};

template <int WHAT, bool HAS_PHI, int K, int STEPS>
__device__ __forceinline__ void cell_corner(const PInv & I, const CWriter & W, const double * ph)
{
	constexpr int NV = phase_comp(WHAT, K, -1);
	if (NV == 0) return;
	double v[NV > 0 ? NV : 1];
	phase_values<WHAT, HAS_PHI, K, true>(I, ph, v);
	#pragma unroll
	for (int s = 0; s < STEPS; s++)
	{
		#pragma unroll
		for (int j = 0; j < NV; j++)
		{
			const double t = __shfl_down_sync(0xffffffffu, v[j], 1 << s);
			if ((1 << s) <= W.after) v[j] += t;
		}
	}
	if (W.write)
	{
		double * p = W.col + acc_index(WHAT, K, 0) * ACC_COLS;
		#pragma unroll
		for (int j = 0; j < NV; j++) p[j * ACC_COLS] += v[j];
	}
}

template <int WHAT, bool HAS_PHI, int STEPS>
__device__ __forceinline__ void cell_corners(const PInv & I, const CWriter & W, const double * ph)
{
	cell_corner<WHAT, HAS_PHI, 0, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 1, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 2, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 3, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 4, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 5, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 6, STEPS>(I, W, ph);
	cell_corner<WHAT, HAS_PHI, 7, STEPS>(I, W, ph);
}

// what the cell at (sx - X, sy - Y, sz - Z) deposited on corner K = 4X + 2Y + Z, i.e. on site (sx, sy, sz); read and cleared
template <int WHAT, int K>
__device__ __forceinline__ void gather_corner(double * acc, int sx, int sy, int sz, double * sum)
{
	constexpr int NV = phase_comp(WHAT, K, -1);
	if (NV == 0) return;
	const int cx = sx - ((K >> 2) & 1), cy = sy - ((K >> 1) & 1), cz = sz - (K & 1);
	if ((unsigned) cx >= GEVB_BX || (unsigned) cy >= GEVB_BY || (unsigned) cz >= UZ) return;
	double * p = acc + acc_index(WHAT, K, 0) * ACC_COLS + ((cz << (GEVB_BX_BITS + GEVB_BY_BITS)) | (cy << GEVB_BX_BITS) | cx);
	#pragma unroll
	for (int j = 0; j < NV; j++)
	{
		sum[phase_comp(WHAT, K, j)] += p[j * ACC_COLS];
		p[j * ACC_COLS] = 0.;
	}
}

// corner and target component of accumulator a (all comparisons are against compile-time constants)
template <int WHAT>
__device__ __forceinline__ void acc_target(int a, int & corner, int & comp)
{
	corner = 0; comp = 0;
	#pragma unroll
	for (int K = 0; K < 8; K++)
	{
		#pragma unroll
		for (int j = 0; j < phase_comp(WHAT, K, -1); j++)
			if (a == acc_index(WHAT, K, j)) { corner = K; comp = phase_comp(WHAT, K, j); }
	}
}

// particle range of unit u = (brick, half)
__device__ __forceinline__ void unit_range(const DParams & D, uint32_t u, uint32_t & first, uint32_t & last)
{
	first = last = 0;
	if (u < 2 * D.G.nbricks) { first = __ldg(D.cell_start + (size_t) u * UCELLS); last = __ldg(D.cell_start + (size_t) (u + 1) * UCELLS); }
}

// slice of the unit's particles that warp w works on: equal shares, a multiple of the warp size
__device__ __forceinline__ void warp_slice(uint32_t first, uint32_t last, int w, uint32_t & wlo, uint32_t & whi)
{
	const uint32_t m = (((last - first + DEP_WARPS - 1) / DEP_WARPS) + 31u) & ~31u;
	wlo = first + (uint32_t) w * m; wlo = wlo < last ? wlo : last;
	whi = wlo + m; whi = whi < last ? whi : last;
}

template <bool HAS_PHI, int THREADS = DEP_THREADS>
__device__ __forceinline__ void stage_unit(const DParams & D, uint32_t unit, uint32_t first, uint32_t last, double * stage)
{
	const BrickGeom & G = D.G;
	double * tphi = stage;
	uint32_t * ctab = (uint32_t *) (stage + UT_SITES + 1);
	if (first == last) return;
	const uint32_t * src = D.cell_start + (size_t) unit * UCELLS;
	for (int k = threadIdx.x; k <= UCELLS; k += THREADS) cp_async4(ctab + k, src + k);
	if (HAS_PHI && threadIdx.x < DX * DY)
	{
		int x0, y0, zl0;
		brick_origin(G, unit >> 1, x0, y0, zl0);
		zl0 += (int) (unit & 1) * UZ;
		const int tx = threadIdx.x % DX, ty = threadIdx.x / DX;
		const size_t gcol = (size_t) wrap_up(y0 + ty, G.N) * G.N + wrap_up(x0 + tx, G.N);
		#pragma unroll
		for (int tz = 0; tz <= UZ; tz++)
		{
			const int plane = zl0 + tz + 1;
			if (plane > G.nzl + 1) break;                       // past the top of the slab: never read
			cp_async8(tphi + (tz * DY + ty) * DX + tx, D.phi + (size_t) plane * G.N * G.N + gcol);
		}
	}
}

template <int WHAT, bool HAS_PHI>
__global__ void __launch_bounds__(DEP_THREADS, 2) k_deposit_cells(DParams D)
{
	constexpr int NCOMP = dep_ncomp(WHAT), NACC = dep_nacc(WHAT);
	extern __shared__ double smem[];
	double * acc = smem;                                    // [NACC][ACC_COLS]
	double * stages = smem + NACC * ACC_COLS;               // [2][USTAGE_DOUBLES]
	int * slotcell = (int *) (stages + 2 * USTAGE_DOUBLES); // [2 * DEP_WARPS] cell of each head / tail column in use, else -1

	const BrickGeom & G = D.G;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int idx = threadIdx.x; idx < NACC * ACC_COLS + 2 * USTAGE_DOUBLES; idx += DEP_THREADS) smem[idx] = 0.;
	if (threadIdx.x < 2 * DEP_WARPS) slotcell[threadIdx.x] = -1;
	__syncthreads();
	const uint32_t nunits = 2 * G.nbricks;
	uint32_t unit = blockIdx.x;
	uint32_t first, last, nfirst, nlast, nnfirst, nnlast;
	unit_range(D, unit, first, last);
	unit_range(D, unit + gridDim.x, nfirst, nlast);
	stage_unit<HAS_PHI>(D, unit, first, last, stages);
	cp_async_commit();
	double pv[6] = {0., 0., 0., 0., 0., 0.};
	{
		uint32_t wlo, whi;
		warp_slice(first, last, w, wlo, whi);
		if (wlo + lane < whi) load_particle(D, wlo + lane, pv);
	}
	int cur = 0;
	while (unit < nunits)
	{
		const uint32_t nunit = unit + gridDim.x;
		unit_range(D, nunit + gridDim.x, nnfirst, nnlast);                   // consumed at the end of this iteration
		if (nunit < nunits) stage_unit<HAS_PHI>(D, nunit, nfirst, nlast, stages + (cur ^ 1) * USTAGE_DOUBLES);
		cp_async_commit();
		uint32_t nwlo, nwhi;                                                 // this warp's slice of the next unit
		warp_slice(nfirst, nlast, w, nwlo, nwhi);
		if (first != last)                                                   // block-uniform
		{
			int x0, y0, zl0;
			brick_origin(G, unit >> 1, x0, y0, zl0);
			zl0 += (int) (unit & 1) * UZ;
			cp_async_wait<1>();                             // everything but the newest group: this unit's stage has landed
			__syncthreads();                                // ... for all threads; also: the previous unit's gather has cleared acc
			const double * tphi = stages + cur * USTAGE_DOUBLES;
			const uint32_t * ctab = (const uint32_t *) (tphi + UT_SITES + 1);
			uint32_t wlo, whi;
			warp_slice(first, last, w, wlo, whi);
			if (lane < 2) slotcell[2 * w + lane] = -1;      // only this warp writes its two entries
			__syncwarp();

			// ---- accumulate: no barrier in here
			for (uint32_t base = wlo; base < whi; base += 32)
			{
				const uint32_t i = base + lane;
				const bool valid = i < whi;
				PInv I;
				CWriter W;
				double phc[8];                              // phi at the eight corners of the particle's cell (index 4X + 2Y + Z)
				int steps;
				{
					int cx = 0, cy = 0, cz = 0, site = 0, cu = 0;
					uint32_t cfirst = i, clast = i + 1;
					if (valid)
					{
						cx = cell_scaled(D, pv[0]); cy = cell_scaled(D, pv[1]); cz = cell_scaled(D, pv[2]);
						// the particle lies in this unit (the storage order is maintained by the re-bin); the masks only keep a broken order memory-safe
						const int sx = (cx - x0) & (GEVB_BX - 1), sy = (cy - y0) & (GEVB_BY - 1), sz = (cz - G.z0 - zl0) & (UZ - 1);
						site = (sz * DY + sy) * DX + sx;
						cu = (sz << (GEVB_BX_BITS + GEVB_BY_BITS)) | (sy << GEVB_BX_BITS) | sx;
						cfirst = ctab[cu]; clast = ctab[cu + 1];
					}
					if (HAS_PHI)
					{
						#pragma unroll
						for (int k = 0; k < 8; k++) phc[k] = tphi[site + corner_offset(k)];
					}
					particle_factors<WHAT, HAS_PHI>(I, D, pv, cx, cy, cz);
					// lanes of this step that share the cell form a contiguous segment [seg_lo, seg_hi)
					const uint32_t step_hi = base + 32 < whi ? base + 32 : whi;
					const uint32_t seg_lo = cfirst > base ? cfirst : base, seg_hi = clast < step_hi ? clast : step_hi;
					const int maxlen = __reduce_max_sync(0xffffffffu, valid ? (int) (seg_hi - seg_lo) : 0);
					steps = maxlen > 1 ? 32 - __clz(maxlen - 1) : 0;
					W.after = valid ? (int) (seg_hi - 1 - i) : 0;
					W.write = valid && i == seg_lo;
					// a cell that reaches out of this warp's slice is accumulated in the warp's head (it began before the slice) or
					// tail column; everything else belongs to this warp alone
					int col = cu;
					if (cfirst < wlo) col = UCELLS + 2 * w; else if (clast > whi) col = UCELLS + 2 * w + 1;
					if (W.write && col >= UCELLS) slotcell[col - UCELLS] = cu;
					W.col = acc + col;
				}
				// request the next step's particle (of this slice, else of this warp's slice of the next unit) while this one is processed
				{
					const uint32_t nb = base + 32;
					const uint32_t j = nb < whi ? nb + lane : nwlo + lane;
					if (j < (nb < whi ? whi : nwhi)) load_particle(D, j, pv);
				}
				// shuffle steps of the segmented sum are warp-uniform: 0 when no two lanes of the step share a cell
				if (steps == 0) cell_corners<WHAT, HAS_PHI, 0>(I, W, phc);
				else if (steps == 1) cell_corners<WHAT, HAS_PHI, 1>(I, W, phc);
				else if (steps == 2) cell_corners<WHAT, HAS_PHI, 2>(I, W, phc);
				else cell_corners<WHAT, HAS_PHI, 5>(I, W, phc);
				__syncwarp();                               // the next step of this warp may add to the same cell
			}
			if (wlo >= whi && nwlo + lane < nwhi) load_particle(D, nwlo + lane, pv);   // idle warp: nothing was prefetched
			__syncthreads();

			// ---- gather + flush: FP64 reductions into HBM, consecutive lanes on consecutive sites of a row; acc is left zeroed
			for (int s = threadIdx.x; s < UT_SITES; s += DEP_THREADS)
			{
				const int sz = s / (DX * DY), r = s - sz * (DX * DY), sy = r / DX, sx = r - sy * DX;
				double sum[NCOMP];
				#pragma unroll
				for (int k = 0; k < NCOMP; k++) sum[k] = 0.;
				gather_corner<WHAT, 0>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 1>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 2>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 3>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 4>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 5>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 6>(acc, sx, sy, sz, sum);
				gather_corner<WHAT, 7>(acc, sx, sy, sz, sum);
				const size_t off = (size_t) (zl0 + sz + 1) * G.N * G.N + (size_t) wrap_up(y0 + sy, G.N) * G.N + wrap_up(x0 + sx, G.N);
				#pragma unroll
				for (int k = 0; k < NCOMP; k++)
					if (sum[k] != 0.) atomicAdd(D.out[k] + off, sum[k]);
			}
			// the head / tail columns in use go to HBM value by value: accumulator a of the column's cell belongs to the site at
			// that cell + its corner offset
			for (int t = threadIdx.x; t < 2 * DEP_WARPS * NACC; t += DEP_THREADS)
			{
				const int slot = t / NACC, a = t - slot * NACC;
				const int c = slotcell[slot];
				if (c < 0) continue;
				double * p = acc + a * ACC_COLS + UCELLS + slot;
				const double v = *p;
				*p = 0.;
				if (v == 0.) continue;
				int K, comp;
				acc_target<WHAT>(a, K, comp);
				const int sx = (c & (GEVB_BX - 1)) + ((K >> 2) & 1), sy = ((c >> GEVB_BX_BITS) & (GEVB_BY - 1)) + ((K >> 1) & 1), sz = (c >> (GEVB_BX_BITS + GEVB_BY_BITS)) + (K & 1);
				const size_t off = (size_t) (zl0 + sz + 1) * G.N * G.N + (size_t) wrap_up(y0 + sy, G.N) * G.N + wrap_up(x0 + sx, G.N);
				atomicAdd(D.out[comp] + off, v);
			}
		}
		else if (nwlo + lane < nwhi) load_particle(D, nwlo + lane, pv);      // empty unit: nothing was prefetched
		unit = nunit; cur ^= 1;
		first = nfirst; last = nlast; nfirst = nnfirst; nlast = nnlast;
	}
}

// =====================================================================================================================
// k_deposit_percell (tuning knob deposit_variant = 2): one thread per CELL, accumulators in registers.
//
// What the profiles of the two kernels above say (profiles/round2_v03): both are bound by the shared-memory data pipe
// (l1tex wavefronts at 71-77 % of peak) and by instruction issue (1455 warp instructions per 32 particles), because every
// particle's 38 contributions are read-added-written in shared memory, shuffled for the segmented sums, and -- in the
// per-cell form -- read and cleared again by the gather.  The reference keeps the contributions of a cell's particles in
// local accumulators (localCube / localEdge, gevolution.hpp:953,1075,1199-1200) and touches the field once per cell; so
// does this kernel: a thread owns a cell, walks the cell's particles (they are contiguous in the sorted order) and
// accumulates in registers -- no shuffle, no shared-memory traffic per particle.  The cell's sums are stored once
// (38 plain stores), a barrier, then the site gather of k_deposit_cells (read-only here) reduces into HBM.
//
//   balance   cells hold different numbers of particles; a warp runs as long as its fullest cell.  The block therefore
//             sorts the unit's 256 cells by particle count first (one shared-memory counting sort over the counts), so the
//             cells of a warp hold about equally many particles; cells above ZC_HEAVY particles are left to whole warps
//             (lanes stride over the cell, one warp reduction per cell).
//   staging   the unit's particles (six contiguous ranges of the SoA arrays) arrive by TMA bulk copies
//             (cp.async.bulk, UBLKCP) that are issued a unit ahead and complete on an mbarrier; units above ZC_PMAX
//             particles are read from global memory directly (their latency is amortised over long loops).
#define ZC_PMAX 320                                    // particles of a unit that the shared-memory stage holds
#define ZC_HEAVY 64                                    // cells with more particles are processed by whole warps
#define ZC_CLASSES 32                                  // count classes of the balance sort (the last one takes 31 .. ZC_HEAVY)
#define ZC_THREADS 192                                 // 2 blocks of 192 threads per SM: 170 registers per thread for the 38 accumulators
#define ZC_PSTRIDE (ZC_PMAX + 2)                       // one staged array (the range is widened to 16-byte boundaries)

// contributions of one particle added to the register accumulators A[] of its cell (corner K)
template <int WHAT, bool HAS_PHI, int K>
__device__ __forceinline__ void accumulate_corner(const PInv & I, const double * phc, double * A)
{
	constexpr int NV = phase_comp(WHAT, K, -1);
	if (NV == 0) return;
	double v[NV > 0 ? NV : 1];
	phase_values<WHAT, HAS_PHI, K, true>(I, phc, v);
	#pragma unroll
	for (int j = 0; j < NV; j++) A[acc_index(WHAT, K, 0) + j] += v[j];
}

template <int WHAT, bool HAS_PHI>
__device__ __forceinline__ void accumulate_particle(const DParams & D, const double * pv, int cx, int cy, int cz, const double * phc, double * A)
{
	PInv I;
	particle_factors<WHAT, HAS_PHI>(I, D, pv, cx, cy, cz);
	accumulate_corner<WHAT, HAS_PHI, 0>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 1>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 2>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 3>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 4>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 5>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 6>(I, phc, A);
	accumulate_corner<WHAT, HAS_PHI, 7>(I, phc, A);
}

// thread 0: request the particles of a unit into the stage (nothing for empty or oversized units); returns whether it did
__device__ __forceinline__ bool request_particles(const DParams & D, uint32_t first, uint32_t last, double * pstage, unsigned long long * bar)
{
	if (first == last || last - first > ZC_PMAX) return false;
	if (threadIdx.x == 0)
	{
		const uint32_t astart = first & ~1u, count = (last - astart + 1u) & ~1u;     // whole 16-byte pieces (the arrays are padded by two elements)
		const double * src[6] = {D.x, D.y, D.z, D.qx, D.qy, D.qz};
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                 // the stage was read through the generic proxy until now
		mbar_expect(bar, 6u * count * (uint32_t) sizeof(double));
		#pragma unroll
		for (int k = 0; k < 6; k++) bulk_g2s(pstage + k * ZC_PSTRIDE, src[k] + astart, count * (uint32_t) sizeof(double), bar);
	}
	return true;
}

// one (x, y) column of tile sites gathers the cells around it: every accumulator of the unit is read exactly once
template <int WHAT, int K>
__device__ __forceinline__ void corner_into(const double * p, double * sum)
{
	constexpr int NV = phase_comp(WHAT, K, -1);
	#pragma unroll
	for (int j = 0; j < NV; j++) sum[phase_comp(WHAT, K, j)] += p[(acc_index(WHAT, K, 0) + j) * UCELLS];
}
template <int WHAT, int X, int Y>
__device__ __forceinline__ void column_from(const double * acc, int sx, int sy, double (* sum)[7])
{
	const int cx = sx - X, cy = sy - Y;
	if ((unsigned) cx >= GEVB_BX || (unsigned) cy >= GEVB_BY) return;
	#pragma unroll
	for (int cz = 0; cz < UZ; cz++)
	{
		const double * p = acc + ((cz << (GEVB_BX_BITS + GEVB_BY_BITS)) | (cy << GEVB_BX_BITS) | cx);
		corner_into<WHAT, 4 * X + 2 * Y>(p, sum[cz]);            // Z = 0: the cell's own plane
		corner_into<WHAT, 4 * X + 2 * Y + 1>(p, sum[cz + 1]);    // Z = 1: the plane above
	}
}

// the particles of one cell accumulated in registers (the thread's cell, or a lane's share of a heavy cell)
template <int WHAT, bool HAS_PHI, bool STAGED>
__device__ __forceinline__ void cell_particles(const DParams & D, const double * pstage, uint32_t base, uint32_t ibegin, uint32_t iend, uint32_t istep,
                                               int cx, int cy, int cz, const double * phc, double * A)
{
	auto fetch = [&](uint32_t i, double * pv)
	{
		if (STAGED)
		{
			const uint32_t j = i - base;
			#pragma unroll
			for (int k = 0; k < 6; k++) pv[k] = pstage[k * ZC_PSTRIDE + j];
		}
		else load_particle(D, i, pv);
	};
	double pv[6];
	if (ibegin < iend) fetch(ibegin, pv);
	for (uint32_t i = ibegin; i < iend; i += istep)
	{
		double cur_pv[6];
		#pragma unroll
		for (int k = 0; k < 6; k++) cur_pv[k] = pv[k];
		if (i + istep < iend) fetch(i + istep, pv);             // the next particle is requested before this one is processed
		accumulate_particle<WHAT, HAS_PHI>(D, cur_pv, cx, cy, cz, phc, A);
	}
}

// counting sort of the unit's cells by particle count, first half: class and rank of this thread's cells
__device__ __forceinline__ void sort_count(const uint32_t * ctab, int * hist, int * nheavy, unsigned char * heavy, int * kc, int * rc)
{
	#pragma unroll
	for (int q = 0; q < 2; q++)
	{
		const int c = threadIdx.x + q * ZC_THREADS;
		kc[q] = rc[q] = 0;
		if (c >= UCELLS) break;
		const uint32_t cnt = ctab[c + 1] - ctab[c];
		const bool is_heavy = cnt > ZC_HEAVY;
		kc[q] = is_heavy ? 0 : (int) (cnt < ZC_CLASSES - 1 ? cnt : ZC_CLASSES - 1);
		rc[q] = atomicAdd(hist + kc[q], 1);
		if (is_heavy) heavy[atomicAdd(nheavy, 1)] = (unsigned char) c;
	}
}
// second half (after a barrier): every warp forms the first slot of each class for itself, then files its cells
__device__ __forceinline__ void sort_place(const int * hist, unsigned char * order, const int * kc, const int * rc)
{
	const int lane = threadIdx.x & 31;
	const int h = hist[lane];
	int s = h;                                              // becomes the number of cells in classes >= lane
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, s, o); if (lane + o < 32) s += t; }
	const int start = s - h;
	#pragma unroll
	for (int q = 0; q < 2; q++)
	{
		const int c = threadIdx.x + q * ZC_THREADS;
		const int first_slot = __shfl_sync(0xffffffffu, start, kc[q]);
		if (c < UCELLS) order[first_slot + rc[q]] = (unsigned char) c;
	}
}

template <int WHAT, bool HAS_PHI>
__global__ void __launch_bounds__(ZC_THREADS, 2) k_deposit_percell(DParams D)
{
	constexpr int NCOMP = dep_ncomp(WHAT), NACC = dep_nacc(WHAT);
	extern __shared__ __align__(16) double smem[];
	double * acc = smem;                                    // [NACC][UCELLS] sums of every cell of the unit
	double * stages = acc + NACC * UCELLS;                  // [2][USTAGE_DOUBLES] phi tile + cell table, one unit ahead
	double * pstage = stages + 2 * USTAGE_DOUBLES;          // [6][ZC_PSTRIDE] the unit's particles
	unsigned long long * bar = (unsigned long long *) (pstage + 6 * ZC_PSTRIDE);
	int * hist = (int *) (bar + 1);                         // [ZC_CLASSES] cells per count class
	int * nheavy = hist + ZC_CLASSES;                       // [2] heavy cells of the unit in work / of the next one
	unsigned char * order = (unsigned char *) (nheavy + 2); // [UCELLS] cells in order of decreasing particle count
	unsigned char * heavy = order + UCELLS;                 // [2][UCELLS] cells above ZC_HEAVY

	const BrickGeom & G = D.G;
	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	for (int idx = threadIdx.x; idx < 2 * USTAGE_DOUBLES; idx += ZC_THREADS) stages[idx] = 0.;
	if (threadIdx.x < ZC_CLASSES) hist[threadIdx.x] = 0;
	if (threadIdx.x == 0) { nheavy[0] = nheavy[1] = 0; mbar_init(bar, 1); }
	__syncthreads();
	const uint32_t nunits = 2 * G.nbricks;
	uint32_t unit = blockIdx.x;
	uint32_t first, last, nfirst, nlast, nnfirst, nnlast;
	unit_range(D, unit, first, last);
	unit_range(D, unit + gridDim.x, nfirst, nlast);
	stage_unit<HAS_PHI, ZC_THREADS>(D, unit, first, last, stages);
	cp_async_commit();
	bool staged = request_particles(D, first, last, pstage, bar);
	uint32_t parity = 0;
	int cur = 0, hb = 0;                                    // pipeline stage in work; heavy-list buffer of the unit in work
	int kc[2], rc[2];
	// the first unit's cells are sorted before the loop; inside the loop the sort of unit k+1 rides behind the gather of unit k
	cp_async_wait<0>();
	__syncthreads();
	if (first != last) sort_count((const uint32_t *) (stages + UT_SITES + 1), hist, nheavy + hb, heavy + hb * UCELLS, kc, rc);
	__syncthreads();
	if (first != last) sort_place(hist, order, kc, rc);
	__syncthreads();
	if (threadIdx.x < ZC_CLASSES) hist[threadIdx.x] = 0;
	while (unit < nunits)
	{
		const uint32_t nunit = unit + gridDim.x;
		unit_range(D, nunit + gridDim.x, nnfirst, nnlast);                   // consumed at the end of this iteration
		if (nunit < nunits) stage_unit<HAS_PHI, ZC_THREADS>(D, nunit, nfirst, nlast, stages + (cur ^ 1) * USTAGE_DOUBLES);
		cp_async_commit();
		int x0 = 0, y0 = 0, zl0 = 0;
		if (first != last)                                                   // block-uniform
		{
			brick_origin(G, unit >> 1, x0, y0, zl0);
			zl0 += (int) (unit & 1) * UZ;
			const double * tphi = stages + cur * USTAGE_DOUBLES;
			const uint32_t * ctab = (const uint32_t *) (tphi + UT_SITES + 1);
			const int numheavy = nheavy[hb];
			const unsigned char * hlist = heavy + hb * UCELLS;
			if (staged) mbar_wait(bar, parity);             // the bulk copies of this unit have completed
			const uint32_t pbase = first & ~1u;

			// ---- one cell per thread (slots in order of decreasing count: the second round holds the emptiest cells)
			for (int slot = threadIdx.x; slot < UCELLS; slot += ZC_THREADS)
			{
				const int c = order[slot];
				const uint32_t cfirst = ctab[c], clast = ctab[c + 1];
				if (clast - cfirst > ZC_HEAVY) continue;
				const int sx = c & (GEVB_BX - 1), sy = (c >> GEVB_BX_BITS) & (GEVB_BY - 1), sz = c >> (GEVB_BX_BITS + GEVB_BY_BITS);
				const int site = (sz * DY + sy) * DX + sx;
				double phc[8];
				if (HAS_PHI)
				{
					#pragma unroll
					for (int k = 0; k < 8; k++) phc[k] = tphi[site + corner_offset(k)];
				}
				double A[NACC];
				#pragma unroll
				for (int a = 0; a < NACC; a++) A[a] = 0.;
				if (staged) cell_particles<WHAT, HAS_PHI, true>(D, pstage, pbase, cfirst, clast, 1, x0 + sx, y0 + sy, G.z0 + zl0 + sz, phc, A);
				else cell_particles<WHAT, HAS_PHI, false>(D, pstage, pbase, cfirst, clast, 1, x0 + sx, y0 + sy, G.z0 + zl0 + sz, phc, A);
				#pragma unroll
				for (int a = 0; a < NACC; a++) acc[a * UCELLS + c] = A[a];
			}
			// ---- cells above ZC_HEAVY: a warp per cell, lanes stride over its particles, one warp reduction per cell
			for (int hc = w; hc < numheavy; hc += ZC_THREADS / 32)
			{
				const int c = hlist[hc];
				const uint32_t cfirst = ctab[c], clast = ctab[c + 1];
				const int sx = c & (GEVB_BX - 1), sy = (c >> GEVB_BX_BITS) & (GEVB_BY - 1), sz = c >> (GEVB_BX_BITS + GEVB_BY_BITS);
				const int site = (sz * DY + sy) * DX + sx;
				double phc[8];
				if (HAS_PHI)
				{
					#pragma unroll
					for (int k = 0; k < 8; k++) phc[k] = tphi[site + corner_offset(k)];
				}
				double A[NACC];
				#pragma unroll
				for (int a = 0; a < NACC; a++) A[a] = 0.;
				if (staged) cell_particles<WHAT, HAS_PHI, true>(D, pstage, pbase, cfirst + lane, clast, 32, x0 + sx, y0 + sy, G.z0 + zl0 + sz, phc, A);
				else cell_particles<WHAT, HAS_PHI, false>(D, pstage, pbase, cfirst + lane, clast, 32, x0 + sx, y0 + sy, G.z0 + zl0 + sz, phc, A);
				#pragma unroll
				for (int a = 0; a < NACC; a++)
				{
					double v = A[a];
					#pragma unroll
					for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
					if (lane == 0) acc[a * UCELLS + c] = v;
				}
			}
		}
		cp_async_wait<0>();                                 // the next unit's phi tile and cell table (this thread's copies)
		__syncthreads();                                    // 1: every cell's sums are in acc; the next stage is visible; the particle stage is free
		if (threadIdx.x == 0) nheavy[hb] = 0;
		if (staged) parity ^= 1;
		staged = request_particles(D, nfirst, nlast, pstage, bar);           // arrives while this unit is gathered
		const bool next_has = nfirst != nlast;
		if (next_has) sort_count((const uint32_t *) (stages + (cur ^ 1) * USTAGE_DOUBLES + UT_SITES + 1), hist, nheavy + (hb ^ 1), heavy + (hb ^ 1) * UCELLS, kc, rc);

		// ---- gather + flush: a thread per (x, y) column of the tile sums the cells around its three sites and issues the
		//      FP64 reductions (RED.ADD.F64) into HBM; consecutive lanes hold consecutive sites of a row
		if (first != last)
			for (int col = threadIdx.x; col < DX * DY; col += ZC_THREADS)
			{
				const int sy = col / DX, sx = col - sy * DX;
				double sum[UZ + 1][7];
				#pragma unroll
				for (int z = 0; z <= UZ; z++)
				{
					#pragma unroll
					for (int k = 0; k < 7; k++) sum[z][k] = 0.;
				}
				column_from<WHAT, 0, 0>(acc, sx, sy, sum);
				column_from<WHAT, 0, 1>(acc, sx, sy, sum);
				column_from<WHAT, 1, 0>(acc, sx, sy, sum);
				column_from<WHAT, 1, 1>(acc, sx, sy, sum);
				const size_t gcol = (size_t) wrap_up(y0 + sy, G.N) * G.N + wrap_up(x0 + sx, G.N);
				#pragma unroll
				for (int z = 0; z <= UZ; z++)
				{
					const size_t off = (size_t) (zl0 + z + 1) * G.N * G.N + gcol;
					#pragma unroll
					for (int k = 0; k < NCOMP; k++)
						if (sum[z][k] != 0.) atomicAdd(D.out[k] + off, sum[z][k]);
				}
			}
		__syncthreads();                                    // 2: the histogram of the next unit is complete; acc is free again
		if (next_has) sort_place(hist, order, kc, rc);
		__syncthreads();                                    // 3: the next unit's slot order is in place
		if (threadIdx.x < ZC_CLASSES) hist[threadIdx.x] = 0;
		unit = nunit; cur ^= 1; hb ^= 1;
		first = nfirst; last = nlast; nfirst = nnfirst; nlast = nnlast;
	}
}

int check_real(const gevb_field * f, int ncomp, const char * who, const char * name)
{
	GEVB_CHECK_ARG(f != NULL, "%s: %s is NULL", who, name);
	GEVB_CHECK_ARG(f->kind == GEVB_REAL, "%s: %s must be a real-space field", who, name);
	GEVB_CHECK_ARG(f->ncomp == ncomp, "%s: %s needs %d components (has %d)", who, name, ncomp, f->ncomp);
	return 0;
}

// tensor maps over the target components: double [nzl + 2][N][N], box = one accumulator tile of a component
template <int WHAT, int FLAGS>
int make_maps(gevb_ctx * c, const DParams & D, DMaps & M)
{
	for (int k = 0; k < dep_ncomp(WHAT); k++) GEVB_TRY(gevb_tensor_map_3d(c, &M.m[k], D.out[k], acc_ax(FLAGS), acc_ay(FLAGS), DZ));
	return 0;
}

template <int WHAT, int FLAGS>
int launch_sites(gevb_ctx * c, const DParams & D, bool has_phi)
{
	const size_t smem = ((size_t) dep_ncomp(WHAT) * acc_sites(FLAGS) + 2 * DEP_STAGE_DOUBLES) * sizeof(double);
	const uint32_t persistent = (uint32_t) c->num_sms * 3;
	const uint32_t grid = D.G.nbricks < persistent ? D.G.nbricks : persistent;
	DMaps M;
	memset(&M, 0, sizeof(M));
	if (FLAGS & DEPF_TMA) GEVB_TRY((make_maps<WHAT, FLAGS>(c, D, M)));
	if (has_phi)
	{
		CUDA_TRY(cudaFuncSetAttribute(k_deposit<WHAT, true, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_deposit<WHAT, true, FLAGS><<<grid, DEP_THREADS, smem, c->stream>>>(D, M);
	}
	else
	{
		CUDA_TRY(cudaFuncSetAttribute(k_deposit<WHAT, false, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_deposit<WHAT, false, FLAGS><<<grid, DEP_THREADS, smem, c->stream>>>(D, M);
	}
	KERNEL_CHECK(c);
	return 0;
}

template <int WHAT>
int launch(gevb_pcls * p, double * const * out, double a, gevb_field * phi, double mass)
{
	gevb_ctx * c = p->ctx;
	if (p->n == 0) return 0;
	const int b = p->cur;
	DParams D;
	D.G = p->geom; D.pow2 = (c->N & (c->N - 1)) == 0;
	D.dx = 1.0 / (double) c->N; D.rN = (double) c->N; D.a = a; D.mass = mass;
	D.cell_start = p->cell_start;
	D.x = p->x[b]; D.y = p->y[b]; D.z = p->z[b]; D.qx = p->qx[b]; D.qy = p->qy[b]; D.qz = p->qz[b];
	D.phi = phi ? phi->data : NULL;
	for (int k = 0; k < 7; k++) D.out[k] = k < dep_ncomp(WHAT) ? out[k] : NULL;
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 2)
	{
		// one thread per cell, register accumulators (k_deposit_percell): two blocks per SM, one unit (half a brick) at a time
		const size_t smem = ((size_t) dep_nacc(WHAT) * UCELLS + 2 * USTAGE_DOUBLES + 6 * ZC_PSTRIDE) * sizeof(double) + 8 + (ZC_CLASSES + 2) * sizeof(int) + 3 * UCELLS;
		const uint32_t persistent = (uint32_t) c->num_sms * 2, nunits = 2 * D.G.nbricks;
		const uint32_t grid = nunits < persistent ? nunits : persistent;
		if (phi)
		{
			CUDA_TRY(cudaFuncSetAttribute(k_deposit_percell<WHAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			k_deposit_percell<WHAT, true><<<grid, ZC_THREADS, smem, c->stream>>>(D);
		}
		else
		{
			CUDA_TRY(cudaFuncSetAttribute(k_deposit_percell<WHAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			k_deposit_percell<WHAT, false><<<grid, ZC_THREADS, smem, c->stream>>>(D);
		}
		KERNEL_CHECK(c);
		return 0;
	}
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 1)
	{
		// per-cell accumulators (k_deposit_cells): two blocks per SM, one unit (half a brick) at a time
		const size_t smem = ((size_t) dep_nacc(WHAT) * ACC_COLS + 2 * USTAGE_DOUBLES) * sizeof(double) + 2 * DEP_WARPS * sizeof(int);
		const uint32_t persistent = (uint32_t) c->num_sms * 2, nunits = 2 * D.G.nbricks;
		const uint32_t grid = nunits < persistent ? nunits : persistent;
		if (phi)
		{
			CUDA_TRY(cudaFuncSetAttribute(k_deposit_cells<WHAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			k_deposit_cells<WHAT, true><<<grid, DEP_THREADS, smem, c->stream>>>(D);
		}
		else
		{
			CUDA_TRY(cudaFuncSetAttribute(k_deposit_cells<WHAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			k_deposit_cells<WHAT, false><<<grid, DEP_THREADS, smem, c->stream>>>(D);
		}
		KERNEL_CHECK(c);
		return 0;
	}
	// the site-tile kernel: deposit_variant 4 flushes with bulk reductions, 0 with one RED per site
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 4) return launch_sites<WHAT, DEPF_BULK>(c, D, phi != NULL);
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 6) return launch_sites<WHAT, DEPF_BULK | DEPF_LOOP>(c, D, phi != NULL);
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 8) return launch_sites<WHAT, DEPF_BULK | DEPF_SPILL>(c, D, phi != NULL);
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 10) return launch_sites<WHAT, DEPF_BULK | DEPF_LOOP | DEPF_SPILL>(c, D, phi != NULL);
	if (gevb_tune(TUNE_DEPOSIT_VARIANT) == 18) return launch_sites<WHAT, DEPF_BULK | DEPF_LOOP | DEPF_SPILL | DEPF_TMA>(c, D, phi != NULL);
	return launch_sites<WHAT, 0>(c, D, phi != NULL);
}

} // namespace

extern "C" int gevb_projection_T00_project(gevb_pcls * p, gevb_field * T00, double a, gevb_field * phi, double coeff)
{
	GEVB_CHECK_ARG(p != NULL, "projection_T00_project: NULL particle handle");
	GEVB_TRY(check_real(T00, 1, "projection_T00_project", "T00"));
	if (phi) GEVB_TRY(check_real(phi, 1, "projection_T00_project", "phi"));
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_T00);
	const double dx = 1.0 / (double) c->N;
	double mass = coeff / (dx * dx * dx); mass *= p->mass; mass /= a;          // gevolution.hpp:945-947
	double * out[7] = {T00->data};
	return launch<DEP_T00>(p, out, a, phi, mass);
}

extern "C" int gevb_projection_Tij_project(gevb_pcls * p, gevb_field * Tij, double a, gevb_field * phi, double coeff)
{
	GEVB_CHECK_ARG(p != NULL, "projection_Tij_project: NULL particle handle");
	GEVB_TRY(check_real(Tij, 6, "projection_Tij_project", "Tij"));
	if (phi) GEVB_TRY(check_real(phi, 1, "projection_Tij_project", "phi"));
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_TIJ);
	const double dx = 1.0 / (double) c->N;
	double mass = coeff / (dx * dx * dx); mass *= p->mass; mass /= a;          // gevolution.hpp:1191-1193
	double * out[7];
	for (int k = 0; k < 6; k++) out[k] = Tij->data + k * Tij->comp_stride;
	return launch<DEP_TIJ>(p, out, a, phi, mass);
}

extern "C" int gevb_projection_T00_Tij_project(gevb_pcls * p, gevb_field * T00, gevb_field * Tij, double a, gevb_field * phi, double coeff)
{
	GEVB_CHECK_ARG(p != NULL, "projection_T00_Tij_project: NULL particle handle");
	GEVB_TRY(check_real(T00, 1, "projection_T00_Tij_project", "T00"));
	GEVB_TRY(check_real(Tij, 6, "projection_T00_Tij_project", "Tij"));
	GEVB_CHECK_ARG(phi != NULL, "projection_T00_Tij_project: phi is required (GR projections)");
	GEVB_TRY(check_real(phi, 1, "projection_T00_Tij_project", "phi"));
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_T00_TIJ);
	const double dx = 1.0 / (double) c->N;
	double mass = coeff / (dx * dx * dx); mass *= p->mass; mass /= a;
	double * out[7] = {T00->data};
	for (int k = 0; k < 6; k++) out[1 + k] = Tij->data + k * Tij->comp_stride;
	return launch<DEP_T00_TIJ>(p, out, a, phi, mass);
}

extern "C" int gevb_scalarProjectionCIC_project(gevb_pcls * p, gevb_field * rho)
{
	GEVB_CHECK_ARG(p != NULL, "scalarProjectionCIC_project: NULL particle handle");
	GEVB_TRY(check_real(rho, 1, "scalarProjectionCIC_project", "rho"));
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_T00);
	const double dx = 1.0 / (double) c->N;
	// plain CIC = T00 projection with e = 1, f = 0 and no 1/a
	double * out[7] = {rho->data};
	return launch<DEP_T00>(p, out, 1.0, NULL, p->mass / (dx * dx * dx));
}

extern "C" int gevb_projection_T0i_project(gevb_pcls * p, gevb_field * T0i, gevb_field * phi, double coeff)
{
	GEVB_CHECK_ARG(p != NULL, "projection_T0i_project: NULL particle handle");
	GEVB_TRY(check_real(T0i, 3, "projection_T0i_project", "T0i"));
	if (phi) GEVB_TRY(check_real(phi, 1, "projection_T0i_project", "phi"));
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_T0I);
	const double dx = 1.0 / (double) c->N;
	double mass = coeff / (dx * dx * dx); mass *= p->mass;                     // gevolution.hpp:1064-1065
	double * out[7];
	for (int k = 0; k < 3; k++) out[k] = T0i->data + k * T0i->comp_stride;
	return launch<DEP_T0I>(p, out, 1.0, phi, mass);
}
