// geodesic.cu -- particle kick and drift
//
//   Particles::updateVel + update_q / update_q_Newton        main.cpp:775; gevolution.hpp:570-678, 709-776
//   Particles::moveParticles + update_pos / update_pos_Newton main.cpp:798; gevolution.hpp:810-871, 900-903
//
// One thread block per brick of 16 x 8 x 4 cells (particles are stored brick by
// brick, gevb_internal.cuh): the block stages the 18 x 10 x 6-site tile (one
// ghost layer below and above) of phi, chi and B_i in shared memory with the
// periodic wrap applied once per tile, then every thread takes particles of
// the brick from the coalesced SoA stream (48 B in, 24 or 48 B out) and gathers
// its ~90 stencil values from shared memory at compile-time offsets.
// The fused kernel does kick and drift in one pass: between main.cpp:775 and
// :798 only `a` changes (rungekutta4bg, :792), positions and fields do not.
// After a drift the new sort key is written and histogrammed (first half of the
// counting sort that re-files the particles); particles that leave the z-slab
// are compacted into send buffers for the two ring neighbours (NCCL P2P).
#include <math.h>
#include <stdlib.h>
#include "gevb_internal.cuh"

namespace {

#define TXO 2                                            // tile column of the brick's first cell: the tile starts at x0 - 2, an even
                                                         // site, because a TMA tensor load wants its box 16-byte aligned in x (x0 - 1 is
                                                         // odd: illegal instruction); column 0 and the last one are never read
#define TX (GEVB_BX + 2 + TXO)
#define TY (GEVB_BY + 2)
#define TZ (GEVB_BZ + 2)
#define TILE_BOX (TX * TY * TZ)                          // 20 x 10 x 6 sites
#define TILE_SITES ((TILE_BOX + 15) / 16 * 16)           // component stride: a component starts on a 128-byte boundary (TMA destination)

struct GParams
{
	BrickGeom G;
	int nranks, pow2;
	size_t plane, csB;
	double dx, rN;             // rN = (double) N: pos/dx == pos*rN exactly when N is a power of two
	double binv_kick, binv_drift;   // 1 / params[1] (the a^2 N that un-scales the stored B, main.cpp:772)
	const double * phi, * chi, * B;
	int nfmax;                 // how many of {phi, chi, B} the tile needs
	int tma;                   // tiles of bricks away from the lattice edge are fetched by the TMA unit (tensor maps in GMaps)
	// kick
	int fn, nf_kick; double dtau_kick, a_kick, bscale_kick;
	// drift
	int nf_drift; double dtau_drift, a_drift, bscale_drift;
	// particles
	int64_t n;
	double * x, * y, * z, * qx, * qy, * qz; int64_t * id; uint32_t * key;
	uint32_t * rank;                     // rank of every particle inside its new cell (NULL: the histogram is only counted, rebin_variant 0)
	const uint32_t * cell_start; uint32_t * cell_count;
	unsigned long long * maxv2;          // bit pattern of the running max of v^2 (>= 0)
	// migration
	unsigned long long * nsend;          // [2]: down, up
	double * sendbuf[2]; int64_t sendcap;
};

struct GMaps { CUtensorMap m[5]; };                      // phi, chi, B0, B1, B2 as 3-D tensors [nzl + 2][N][N], box = one tile

// scaled coordinate pos/dx (LATfield2 drivers use pos/dx; the product is bit-identical for power-of-two N)
__device__ __forceinline__ double scaled(const GParams & P, double p) { return P.pow2 ? p * P.rN : p / P.dx; }
__device__ __forceinline__ double by_dx(const GParams & P, double v) { return P.pow2 ? v * P.rN : v / P.dx; }
__device__ __forceinline__ int cell_scaled(double s, int N)
{
	int c = (int) floor(s);
	c = c >= N ? N - 1 : c;
	return c < 0 ? 0 : c;
}

// T(f, i, j, k): value of tile component f at the particle's cell + (i, j, k), i, j, k in {-1, 0, 1}
#define T(f, i, j, k) t[(f) * TILE_SITES + (k) * (TX * TY) + (j) * TX + (i)]

// one-sided CIC gradient, gevolution.hpp:585-596 (GRADIENT_ORDER == 1)
__device__ __forceinline__ void grad_cic(const double * t, int f, const double * r, double * g)
{
	const double f000 = T(f, 0, 0, 0), f100 = T(f, 1, 0, 0), f010 = T(f, 0, 1, 0), f110 = T(f, 1, 1, 0);
	const double f001 = T(f, 0, 0, 1), f101 = T(f, 1, 0, 1), f011 = T(f, 0, 1, 1), f111 = T(f, 1, 1, 1);
	g[0] = (1. - r[1]) * (1. - r[2]) * (f100 - f000);
	g[1] = (1. - r[0]) * (1. - r[2]) * (f010 - f000);
	g[2] = (1. - r[0]) * (1. - r[1]) * (f001 - f000);
	g[0] += r[1] * (1. - r[2]) * (f110 - f010);
	g[1] += r[0] * (1. - r[2]) * (f110 - f100);
	g[2] += r[0] * (1. - r[1]) * (f101 - f100);
	g[0] += (1. - r[1]) * r[2] * (f101 - f001);
	g[1] += (1. - r[0]) * r[2] * (f011 - f001);
	g[2] += (1. - r[0]) * r[1] * (f011 - f010);
	g[0] += r[1] * r[2] * (f111 - f011);
	g[1] += r[0] * r[2] * (f111 - f101);
	g[2] += r[0] * r[1] * (f111 - f110);
}

// trilinear interpolation, gevolution.hpp:820-827
__device__ __forceinline__ double tri_cic(const double * t, int f, const double * r)
{
	double v = T(f, 0, 0, 0) * (1. - r[0]) * (1. - r[1]) * (1. - r[2]);
	v += T(f, 1, 0, 0) * r[0] * (1. - r[1]) * (1. - r[2]);
	v += T(f, 0, 1, 0) * (1. - r[0]) * r[1] * (1. - r[2]);
	v += T(f, 1, 1, 0) * r[0] * r[1] * (1. - r[2]);
	v += T(f, 0, 0, 1) * (1. - r[0]) * (1. - r[1]) * r[2];
	v += T(f, 1, 0, 1) * r[0] * (1. - r[1]) * r[2];
	v += T(f, 0, 1, 1) * (1. - r[0]) * r[1] * r[2];
	v += T(f, 1, 1, 1) * r[0] * r[1] * r[2];
	return v;
}

// update_q (gevolution.hpp:570-678) / update_q_Newton (:709-776); returns q^2 after the kick
__device__ __forceinline__ double kick(const GParams & P, const double * t, const double * r, double * q, int64_t pid)
{
	double g[3], v2;
	if (P.fn == GEVB_INITIALIZE_Q_IC_BASIC)
	{
		// initialize_q_ic_basic (ic_basic.hpp:118-152): q = -coeff grad(potential) / dx; with two potentials every eighth ID uses the second
		const int f = (P.nf_kick > 1 && (pid & 7) == 0) ? 1 : 0;                   // :124-127
		grad_cic(t, f, r, g);                                                      // :129-140
		v2 = 0.;
		#pragma unroll
		for (int i = 0; i < 3; i++) { q[i] = -by_dx(P, g[i]) * P.dtau_kick; v2 += q[i] * q[i]; }   // :142-150
		return v2;
	}
	if (P.fn == GEVB_UPDATE_Q)
	{
		v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                             // :581
		double e2 = v2 + P.a_kick * P.a_kick;                                      // :582
		grad_cic(t, 0, r, g);                                                      // :585-596
		// no division and no square root in the kick: r = e2^(-1/2) gives 1/e2 = r r and sqrt(e2) = e2 r (each within 2 ulp)
		const double rs = rsqrt(e2);
		const double inv_e2 = rs * rs;
		const double boost = (v2 + e2) * inv_e2;
		g[0] *= boost; g[1] *= boost; g[2] *= boost;                               // :613-615
		if (P.nf_kick >= 2)
		{
			double gc[3]; grad_cic(t, 1, r, gc);                                   // :617-631
			g[0] -= gc[0]; g[1] -= gc[1]; g[2] -= gc[2];
		}
		e2 = e2 * rs;                                                              // :633 sqrt(e2)
		if (P.nf_kick >= 3)
		{
			double pg0, pg1, pg2;
			// :637-642
			pg0 = ((1. - r[2]) * (T(3, 1, 0, 0) - T(3, 0, 0, 0)) + r[2] * (T(3, 1, 0, 1) - T(3, 0, 0, 1))) * q[1];
			pg0 += ((1. - r[1]) * (T(4, 1, 0, 0) - T(4, 0, 0, 0)) + r[1] * (T(4, 1, 1, 0) - T(4, 0, 1, 0))) * q[2];
			pg0 += (1. - r[1]) * (1. - r[2]) * ((r[0] - 1.) * T(2, -1, 0, 0) + (1. - 2. * r[0]) * T(2, 0, 0, 0) + r[0] * T(2, 1, 0, 0)) * q[0];
			pg0 += r[1] * (1. - r[2]) * ((r[0] - 1.) * T(2, -1, 1, 0) + (1. - 2. * r[0]) * T(2, 0, 1, 0) + r[0] * T(2, 1, 1, 0)) * q[0];
			pg0 += (1. - r[1]) * r[2] * ((r[0] - 1.) * T(2, -1, 0, 1) + (1. - 2. * r[0]) * T(2, 0, 0, 1) + r[0] * T(2, 1, 0, 1)) * q[0];
			pg0 += r[1] * r[2] * ((r[0] - 1.) * T(2, -1, 1, 1) + (1. - 2. * r[0]) * T(2, 0, 1, 1) + r[0] * T(2, 1, 1, 1)) * q[0];
			// :644-649
			pg1 = ((1. - r[0]) * (T(4, 0, 1, 0) - T(4, 0, 0, 0)) + r[0] * (T(4, 1, 1, 0) - T(4, 1, 0, 0))) * q[2];
			pg1 += ((1. - r[2]) * (T(2, 0, 1, 0) - T(2, 0, 0, 0)) + r[2] * (T(2, 0, 1, 1) - T(2, 0, 0, 1))) * q[0];
			pg1 += (1. - r[0]) * (1. - r[2]) * ((r[1] - 1.) * T(3, 0, -1, 0) + (1. - 2. * r[1]) * T(3, 0, 0, 0) + r[1] * T(3, 0, 1, 0)) * q[1];
			pg1 += r[0] * (1. - r[2]) * ((r[1] - 1.) * T(3, 1, -1, 0) + (1. - 2. * r[1]) * T(3, 1, 0, 0) + r[1] * T(3, 1, 1, 0)) * q[1];
			pg1 += (1. - r[0]) * r[2] * ((r[1] - 1.) * T(3, 0, -1, 1) + (1. - 2. * r[1]) * T(3, 0, 0, 1) + r[1] * T(3, 0, 1, 1)) * q[1];
			pg1 += r[0] * r[2] * ((r[1] - 1.) * T(3, 1, -1, 1) + (1. - 2. * r[1]) * T(3, 1, 0, 1) + r[1] * T(3, 1, 1, 1)) * q[1];
			// :651-656
			pg2 = ((1. - r[1]) * (T(2, 0, 0, 1) - T(2, 0, 0, 0)) + r[1] * (T(2, 0, 1, 1) - T(2, 0, 1, 0))) * q[0];
			pg2 += ((1. - r[0]) * (T(3, 0, 0, 1) - T(3, 0, 0, 0)) + r[0] * (T(3, 1, 0, 1) - T(3, 1, 0, 0))) * q[1];
			pg2 += (1. - r[0]) * (1. - r[1]) * ((r[2] - 1.) * T(4, 0, 0, -1) + (1. - 2. * r[2]) * T(4, 0, 0, 0) + r[2] * T(4, 0, 0, 1)) * q[2];
			pg2 += r[0] * (1. - r[1]) * ((r[2] - 1.) * T(4, 1, 0, -1) + (1. - 2. * r[2]) * T(4, 1, 0, 0) + r[2] * T(4, 1, 0, 1)) * q[2];
			pg2 += (1. - r[0]) * r[1] * ((r[2] - 1.) * T(4, 0, 1, -1) + (1. - 2. * r[2]) * T(4, 0, 1, 0) + r[2] * T(4, 0, 1, 1)) * q[2];
			pg2 += r[0] * r[1] * ((r[2] - 1.) * T(4, 1, 1, -1) + (1. - 2. * r[2]) * T(4, 1, 1, 0) + r[2] * T(4, 1, 1, 1)) * q[2];
			const double s = P.binv_kick * e2 * inv_e2;                            // :658-660 (pg / params[1] / e2, e2 now holds e)
			g[0] += pg0 * s;
			g[1] += pg1 * s;
			g[2] += pg2 * s;
		}
		v2 = 0.;
		#pragma unroll
		for (int i = 0; i < 3; i++) { q[i] -= by_dx(P, P.dtau_kick * e2 * g[i]); v2 += q[i] * q[i]; }   // :664-668
	}
	else
	{
		grad_cic(t, 0, r, g);                                                      // :719-730
		if (P.nf_kick >= 2)
		{
			double gc[3]; grad_cic(t, 1, r, gc);                                   // :747-761
			g[0] -= gc[0]; g[1] -= gc[1]; g[2] -= gc[2];
		}
		v2 = 0.;
		#pragma unroll
		for (int i = 0; i < 3; i++) { q[i] -= by_dx(P, P.dtau_kick * P.a_kick * g[i]); v2 += q[i] * q[i]; }   // :764-768
	}
	return v2;                                                                     // :670 / :770 return v2/a/a: the monotonic division is applied to the maximum on the host
}

// update_pos (gevolution.hpp:810-871) / update_pos_Newton (:900-903)
// returns the squared displacement for displace_pcls_ic_basic (its reduction output, ic_basic.hpp:88-89), else 0
__device__ __forceinline__ double drift(const GParams & P, const double * t, const double * r, const double * q, double * pos, int64_t pid)
{
	if (P.fn == GEVB_DISPLACE_PCLS_IC_BASIC)
	{
		// displace_pcls_ic_basic (ic_basic.hpp:60-92): pos += coeff grad(xi) / dx
		double g[3], d2 = 0.;
		const int f = (P.nf_drift > 1 && (pid & 7) == 0) ? 1 : 0;                  // :65-68
		grad_cic(t, f, r, g);                                                      // :70-81
		#pragma unroll
		for (int l = 0; l < 3; l++) { g[l] = by_dx(P, g[l]); d2 += g[l] * g[l]; pos[l] += P.dtau_drift * g[l]; }   // :83-91
		return P.dtau_drift * P.dtau_drift * d2;
	}
	if (P.fn != GEVB_UPDATE_Q)
	{
		#pragma unroll
		for (int l = 0; l < 3; l++) pos[l] += P.dtau_drift * q[l] / P.a_drift;     // :902
		return 0.;
	}
	double v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                           // :813
	const double e2 = v2 + P.a_drift * P.a_drift;                                  // :814
	double ph = 0., ch = 0.;
	if (P.nf_drift >= 1) ph = tri_cic(t, 0, r);                                    // :820-827
	if (P.nf_drift >= 2) ch = tri_cic(t, 1, r);                                    // :832-839
	const double rs = rsqrt(e2);                                                   // 1/e2 = rs rs, 1/sqrt(e2) = rs
	v2 = (1. + (3. - v2 * (rs * rs)) * ph - ch) * rs;                              // :842 (... / sqrt(e2))
	double v[3] = {q[0] * v2, q[1] * v2, q[2] * v2};                               // :844-846
	if (P.nf_drift >= 3)
	{
		double b[3];
		b[0] = T(2, 0, 0, 0) * (1. - r[1]) * (1. - r[2]);                          // :852
		b[1] = T(3, 0, 0, 0) * (1. - r[0]) * (1. - r[2]);                          // :853
		b[2] = T(4, 0, 0, 0) * (1. - r[0]) * (1. - r[1]);                          // :854
		b[1] += T(3, 1, 0, 0) * r[0] * (1. - r[2]);                                // :855
		b[2] += T(4, 1, 0, 0) * r[0] * (1. - r[1]);                                // :856
		b[0] += T(2, 0, 1, 0) * r[1] * (1. - r[2]);                                // :857
		b[2] += T(4, 0, 1, 0) * (1. - r[0]) * r[1];                                // :858
		b[0] += T(2, 0, 0, 1) * (1. - r[1]) * r[2];                                // :859
		b[1] += T(3, 0, 0, 1) * (1. - r[0]) * r[2];                                // :860
		b[1] += T(3, 1, 0, 1) * r[0] * r[2];                                       // :861
		b[0] += T(2, 0, 1, 1) * r[1] * r[2];                                       // :862
		b[2] += T(4, 1, 1, 0) * r[0] * r[1];                                       // :863
		#pragma unroll
		for (int l = 0; l < 3; l++) pos[l] += P.dtau_drift * (v[l] + b[l] * P.binv_drift);     // :865 (b / params[1])
	}
	else
	{
		#pragma unroll
		for (int l = 0; l < 3; l++) pos[l] += P.dtau_drift * v[l];                 // :869
	}
	return 0.;
}
#undef T

// periodic wrap into [0,1): p - floor(p), a result that rounds to 1 maps to 0 (edge semantics)
__device__ __forceinline__ double wrap_pos(double p)
{
	double w = p - floor(p);
	return w >= 1.0 ? 0. : w;
}

__device__ __forceinline__ int wrap_index(int v, int N)
{
	// v is within one lattice length of [0, N) except for bricks of a partial super-brick (never staged with particles)
	if (v < 0) v += N; else if (v >= N) v -= N;
	return (unsigned) v >= (unsigned) N ? ((v % N) + N) % N : v;
}

__device__ __forceinline__ void cp_async8(double * smem_dst, const double * gmem_src)
{
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned) __cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long * bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned long long * bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long * bar, uint32_t parity)
{
	uint32_t done = 0;
	while (!done)
		asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA tensor load: the box of the tensor map at (x, y, z) -> dense box at smem_dst (128-byte aligned), completion on the mbarrier
__device__ __forceinline__ void tensor_load_3d(double * smem_dst, const CUtensorMap * map, int x, int y, int z, unsigned long long * bar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		:: "r"(smem_u32(smem_dst)), "l"((unsigned long long) map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}

// does the tile of the brick at (x0, y0) lie inside the lattice in x and y (no periodic wrap)?  Planes past the top of the slab
// are out of the tensor's bounds and arrive as zeros; they belong to a partial brick and are never read.
__device__ __forceinline__ bool tile_inside(const BrickGeom & G, int x0, int y0)
{
	return x0 >= TXO && x0 - TXO + TX <= G.N && y0 >= 1 && y0 + GEVB_BY + 1 <= G.N;
}

// the same tile by tensor loads of the TMA unit (UTMALDG): one instruction per component, issued by one thread
__device__ __forceinline__ void stage_tile_tma(const GParams & P, const GMaps & maps, int x0, int y0, int zl0, double * tile, unsigned long long * bar)
{
	if (threadIdx.x != 0) return;
	const int ncomp = P.nfmax >= 3 ? 5 : P.nfmax;
	mbar_expect(bar, (uint32_t) (ncomp * TILE_BOX * sizeof(double)));
	for (int f = 0; f < ncomp; f++) tensor_load_3d(tile + f * TILE_SITES, &maps.m[f], x0 - TXO, y0 - 1, zl0, bar);
}

// asynchronous copy (LDGSTS) of one brick's field tile into shared memory: site (lx, ly, lz) of the tile is
// lattice site (x0 - TXO + lx, y0 - 1 + ly, local z = zl0 - 1 + lz); a thread keeps its (lx, ly) column and
// walks z, so the periodic wrap is resolved once per thread and all copies of a thread are in flight together
template <int THREADS>
__device__ __forceinline__ void stage_tile(const GParams & P, uint32_t brick, double * tile)
{
	const BrickGeom & G = P.G;
	int x0, y0, zl0;
	brick_origin(G, brick, x0, y0, zl0);
	for (int col_id = threadIdx.x; col_id < (GEVB_BX + 2) * TY; col_id += THREADS)
	{
	const int lx = col_id % (GEVB_BX + 2) + TXO - 1, ly = col_id / (GEVB_BX + 2);
	const size_t col = (size_t) wrap_index(y0 - 1 + ly, G.N) * G.N + wrap_index(x0 - TXO + lx, G.N);
	const int ncomp = P.nfmax >= 3 ? 5 : P.nfmax;
	#pragma unroll
	for (int lz = 0; lz < TZ; lz++)
	{
		const int plane = zl0 + lz;                              // ghost-offset plane index of local z = zl0 - 1 + lz
		if (plane > G.nzl + 1) break;                            // partial brick at the top of the slab: never read
		const size_t off = (size_t) plane * P.plane + col;
		double * d = tile + (lz * TY + ly) * TX + lx;
		if (ncomp >= 1) cp_async8(d, P.phi + off);
		if (ncomp >= 2) cp_async8(d + TILE_SITES, P.chi + off);
		if (ncomp >= 5)
		{
			cp_async8(d + 2 * TILE_SITES, P.B + off);
			cp_async8(d + 3 * TILE_SITES, P.B + P.csB + off);
			cp_async8(d + 4 * TILE_SITES, P.B + 2 * P.csB + off);
		}
	}
	}
}

__device__ __forceinline__ void brick_range(const GParams & P, uint32_t b, uint32_t & first, uint32_t & last)
{
	first = last = 0;
	if (b < P.G.nbricks) { first = __ldg(P.cell_start + (size_t) b * GEVB_BRICK_CELLS); last = __ldg(P.cell_start + (size_t) (b + 1) * GEVB_BRICK_CELLS); }
}

// MODE 0: kick only, 1: drift only, 2: fused kick + drift.
// Persistent blocks walk the bricks with stride gridDim.x.  NBUF == 2: the field tile of the next brick is copied
// asynchronously into the second shared-memory buffer while the particles of the current one are processed
// (2 blocks of 256 threads per SM).  NBUF == 1: one tile buffer per block and twice as many, smaller blocks per SM --
// a block that waits for its tile leaves the SM to the others.  In both forms the first particle of the next brick is
// requested before the current brick's closing barrier, so no DRAM latency is exposed at a brick boundary.
// PREF 1: the next particle of every thread (the next of this brick, else the first of the next brick) is fetched by
// asynchronous copies into a private shared-memory slot while the current one is processed -- a register prefetch at
// the bottom of the loop is consumed immediately at its top and hides nothing within the warp.  PREF 2: the same
// through a second set of registers, loaded at the top of the iteration.
template <int MODE, int THREADS, int NBUF, int MINB, int PREF>
__global__ void __launch_bounds__(THREADS, MINB) k_geodesic(GParams P, const __grid_constant__ GMaps maps)
{
	extern __shared__ __align__(128) double smem[];
	const BrickGeom & G = P.G;
	const int ncomp = P.nfmax >= 3 ? 5 : (P.nfmax > 0 ? P.nfmax : 1);
	const int tile_doubles = ncomp * TILE_SITES;
	double * slot = smem + NBUF * tile_doubles + threadIdx.x;           // [6][THREADS] staging of the prefetched particle (PREF)
	// TMA staging (single tile buffer only): completion barrier behind everything else; `landed` = parity of the next completion
	unsigned long long * bar = (unsigned long long *) (smem + NBUF * tile_doubles + (PREF == 1 ? 6 * THREADS : 0));
	const bool tma = NBUF == 1 && P.tma;
	uint32_t landed = 0;
	bool by_tma = false;                                                // how the current brick's tile was requested
	if (tma) { if (threadIdx.x == 0) mbar_init(bar, 1); __syncthreads(); }
	uint32_t first, last, nfirst, nlast, nnfirst, nnlast;
	uint32_t brick = blockIdx.x;
	brick_range(P, brick, first, last);
	brick_range(P, brick + gridDim.x, nfirst, nlast);
	if (first != last)
	{
		int x0, y0, zl0;
		brick_origin(G, brick, x0, y0, zl0);
		by_tma = tma && tile_inside(G, x0, y0);
		if (by_tma) stage_tile_tma(P, maps, x0, y0, zl0, smem, bar); else stage_tile<THREADS>(P, brick, smem);
	}
	cp_async_commit();
	int cur = 0;
	double vmax = 0.;
	uint32_t pend_i = 0xffffffffu, pend_rank = 0;                       // rank returned by the histogram's atomic, stored one particle later
	uint32_t i = first + threadIdx.x;
	double pos[3] = {0., 0., 0.}, q[3] = {0., 0., 0.};
	if (i < last) { pos[0] = P.x[i]; pos[1] = P.y[i]; pos[2] = P.z[i]; q[0] = P.qx[i]; q[1] = P.qy[i]; q[2] = P.qz[i]; }
	while (brick < G.nbricks)
	{
		const uint32_t nbrick = brick + gridDim.x;
		bool have_next = false;                                          // registers already hold this thread's first particle of the next brick
		brick_range(P, nbrick + gridDim.x, nnfirst, nnlast);             // consumed at the end of this iteration
		if (NBUF == 2)
		{
			if (nfirst != nlast) stage_tile<THREADS>(P, nbrick, smem + (cur ^ 1) * tile_doubles);
			cp_async_commit();
		}
		if (first != last)                                               // block-uniform
		{
			int x0, y0, zl0;
			brick_origin(G, brick, x0, y0, zl0);
			if (by_tma) { mbar_wait(bar, landed); landed ^= 1; }         // this brick's tile has landed (TMA: the barrier's completion makes it visible)
			else
			{
				if (NBUF == 2) cp_async_wait<1>(); else cp_async_wait<0>();
				__syncthreads();
			}
			const double * tile = smem + (NBUF == 2 ? cur * tile_doubles : 0);
			while (i < last)
			{
				uint32_t inext = i + THREADS;
				bool fetched = false;
				double pn[3] = {0., 0., 0.}, qn[3] = {0., 0., 0.};
				if (PREF == 2)
				{
					const uint32_t j = inext < last ? inext : nfirst + threadIdx.x;
					fetched = inext < last || j < nlast;
					if (fetched) { pn[0] = __ldcv(P.x + j); pn[1] = __ldcv(P.y + j); pn[2] = __ldcv(P.z + j); qn[0] = __ldcv(P.qx + j); qn[1] = __ldcv(P.qy + j); qn[2] = __ldcv(P.qz + j); }
				}
				if (PREF == 1)
				{
					// next particle of this brick, else this thread's first particle of the next brick
					const uint32_t j = inext < last ? inext : nfirst + threadIdx.x;
					fetched = inext < last || j < nlast;
					if (fetched)
					{
						cp_async8(slot, P.x + j); cp_async8(slot + THREADS, P.y + j); cp_async8(slot + 2 * THREADS, P.z + j);
						cp_async8(slot + 3 * THREADS, P.qx + j); cp_async8(slot + 4 * THREADS, P.qy + j); cp_async8(slot + 5 * THREADS, P.qz + j);
					}
					cp_async_commit();
				}
				// the previous particle's rank has had a whole iteration to come back from L2
				if ((MODE == 1 || MODE == 2) && pend_i != 0xffffffffu) { P.rank[pend_i] = pend_rank; pend_i = 0xffffffffu; }
				// the particle's cell and ref_dist = frac(pos/dx) (LATfield2 updateVel / moveParticles drivers)
				double r[3];
				const double * t;
				{
					const double sx = scaled(P, pos[0]), sy = scaled(P, pos[1]), sz = scaled(P, pos[2]);
					const int cx = cell_scaled(sx, G.N), cy = cell_scaled(sy, G.N), cz = cell_scaled(sz, G.N);
					r[0] = sx - floor(sx); r[1] = sy - floor(sy); r[2] = sz - floor(sz);      // modf(pos/dx) for pos >= 0
					t = tile + ((cz - G.z0 - zl0 + 1) * TY + (cy - y0 + 1)) * TX + (cx - x0 + TXO);
				}
				const int64_t pid = P.fn >= GEVB_DISPLACE_PCLS_IC_BASIC ? P.id[i] : 0;      // only the IC callbacks look at the ID
				if (MODE == 0 || MODE == 2)
				{
					const double v2 = kick(P, t, r, q, pid);
					vmax = fmax(vmax, v2);
					P.qx[i] = q[0]; P.qy[i] = q[1]; P.qz[i] = q[2];
				}
				if (MODE == 1 || MODE == 2)
				{
					const double d2 = drift(P, t, r, q, pos, pid);
					if (MODE == 1) vmax = fmax(vmax, d2);
					pos[0] = wrap_pos(pos[0]); pos[1] = wrap_pos(pos[1]); pos[2] = wrap_pos(pos[2]);
					const int cx = cell_scaled(scaled(P, pos[0]), G.N), cy = cell_scaled(scaled(P, pos[1]), G.N), cz = cell_scaled(scaled(P, pos[2]), G.N);
					const int zl = cz - G.z0;
					uint32_t key;
					if (P.nranks > 1 && (zl < 0 || zl >= G.nzl))
					{
						// at most one slab per move (main.cpp:281-286): periodic distance decides the neighbour
						const int d = (zl + G.N) % G.N;
						const int dir = d < G.N / 2 ? 1 : 0;
						const unsigned long long slot = atomicAdd(P.nsend + dir, 1ull);
						if ((int64_t) slot < P.sendcap)
						{
							double * sb = P.sendbuf[dir];
							sb[slot] = pos[0]; sb[P.sendcap + slot] = pos[1]; sb[2 * P.sendcap + slot] = pos[2];
							sb[3 * P.sendcap + slot] = q[0]; sb[4 * P.sendcap + slot] = q[1]; sb[5 * P.sendcap + slot] = q[2];
							sb[6 * P.sendcap + slot] = __longlong_as_double((long long) P.id[i]);
						}
						key = GEVB_INVALID_KEY;
					}
					else
					{
						key = brick_key(G, cx, cy, zl);
						// histogram of the counting sort (particles.cu); what the atomic returns is the particle's slot inside its new cell
						if (P.rank) { pend_rank = atomicAdd(P.cell_count + key, 1u); pend_i = i; }
						else atomicAdd(P.cell_count + key, 1u);
					}
					P.x[i] = pos[0]; P.y[i] = pos[1]; P.z[i] = pos[2];
					P.key[i] = key;
				}
				if (PREF)
				{
					// the copies were issued at the top of this iteration; only this thread reads its slot, so no barrier
					if (PREF == 1) cp_async_wait<0>();
					i = inext;
					if (fetched)
					{
						if (PREF == 1) { pos[0] = slot[0]; pos[1] = slot[THREADS]; pos[2] = slot[2 * THREADS]; q[0] = slot[3 * THREADS]; q[1] = slot[4 * THREADS]; q[2] = slot[5 * THREADS]; }
						else { pos[0] = pn[0]; pos[1] = pn[1]; pos[2] = pn[2]; q[0] = qn[0]; q[1] = qn[1]; q[2] = qn[2]; }
					}
					have_next = fetched && i >= last;
					if (have_next) i = nfirst + threadIdx.x;
				}
				else
				{
					i += THREADS;
					if (i < last) { pos[0] = P.x[i]; pos[1] = P.y[i]; pos[2] = P.z[i]; q[0] = P.qx[i]; q[1] = P.qy[i]; q[2] = P.qz[i]; }
				}
				if (have_next) break;
			}
		}
		// request the first particle of the next brick (unless it is already in the registers), then close this one
		if (!have_next)
		{
			i = nfirst + threadIdx.x;
			if (i < nlast) { pos[0] = P.x[i]; pos[1] = P.y[i]; pos[2] = P.z[i]; q[0] = P.qx[i]; q[1] = P.qy[i]; q[2] = P.qz[i]; }
		}
		if (first != last) __syncthreads();                      // the tile is free again
		if (NBUF == 1)
		{
			if (nfirst != nlast)
			{
				int x0, y0, zl0;
				brick_origin(G, nbrick, x0, y0, zl0);
				by_tma = tma && tile_inside(G, x0, y0);
				if (by_tma) stage_tile_tma(P, maps, x0, y0, zl0, smem, bar); else stage_tile<THREADS>(P, nbrick, smem);
			}
			cp_async_commit();
		}
		brick = nbrick; first = nfirst; last = nlast; nfirst = nnfirst; nlast = nnlast; cur ^= 1;
	}
	if ((MODE == 1 || MODE == 2) && pend_i != 0xffffffffu) P.rank[pend_i] = pend_rank;
	{
		for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_down_sync(0xffffffffu, vmax, o));
		if ((threadIdx.x & 31) == 0 && vmax > 0.) atomicMax(P.maxv2, (unsigned long long) __double_as_longlong(vmax));
	}
}

// received particles (7 x cap SoA staging) appended behind the live ones, keys computed and histogrammed.
// cnt != NULL (peer-memory migration): the number received is cnt[which], the records go behind the cnt[0] received
// before them when which == 1, never beyond `room` slots of the arrays (lost[2] counts what did not fit)
__global__ void k_append_received(int64_t nrecv, const double * __restrict__ rb, int64_t cap, int64_t at,
                                  double * x, double * y, double * z, double * qx, double * qy, double * qz, int64_t * id, uint32_t * key,
                                  BrickGeom G, double dx, uint32_t * cell_count, unsigned long long * lost, uint32_t * rank,
                                  const unsigned long long * cnt = NULL, int which = 0, int64_t room = 0)
{
	if (cnt)
	{
		nrecv = (int64_t) cnt[which];
		if (nrecv > cap) nrecv = cap;
		if (which == 1) at += (int64_t) (cnt[0] < (unsigned long long) cap ? cnt[0] : (unsigned long long) cap);
		if (at + nrecv > room)
		{
			if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(lost + 2, (unsigned long long) (at + nrecv - room));
			nrecv = room > at ? room - at : 0;
		}
	}
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < nrecv; i += (int64_t) gridDim.x * blockDim.x)
	{
		const double px = rb[i], py = rb[cap + i], pz = rb[2 * cap + i];
		x[at + i] = px; y[at + i] = py; z[at + i] = pz;
		qx[at + i] = rb[3 * cap + i]; qy[at + i] = rb[4 * cap + i]; qz[at + i] = rb[5 * cap + i];
		id[at + i] = (int64_t) __double_as_longlong(rb[6 * cap + i]);
		const int cx = cell_of(px, dx, G.N), cy = cell_of(py, dx, G.N);
		int cz = cell_of(pz, dx, G.N) - G.z0;
		if (cz < 0 || cz >= G.nzl) { atomicAdd(lost, 1ull); cz = cz < 0 ? 0 : G.nzl - 1; }   // moved farther than one slab: reported by the host
		const uint32_t k = brick_key(G, cx, cy, cz);
		key[at + i] = k;
		const uint32_t r = atomicAdd(cell_count + k, 1u);
		if (rank) rank[at + i] = r;
	}
}

int check_fields(gevb_ctx * c, gevb_field * const * fields, int nfields, const char * who)
{
	GEVB_CHECK_ARG(nfields >= 0 && nfields <= 3, "%s: nfields must be 0..3 (got %d)", who, nfields);
	GEVB_CHECK_ARG(nfields == 0 || fields != NULL, "%s: fields is NULL", who);
	for (int i = 0; i < nfields; i++)
	{
		GEVB_CHECK_ARG(fields[i] != NULL, "%s: fields[%d] is NULL", who, i);
		GEVB_CHECK_ARG(fields[i]->kind == GEVB_REAL && fields[i]->ctx == c, "%s: fields[%d] must be a real field of the same context", who, i);
		GEVB_CHECK_ARG(fields[i]->ncomp == (i == 2 ? 3 : 1), "%s: fields[%d] has %d components", who, i, fields[i]->ncomp);
	}
	return 0;
}

void base_params(GParams & P, gevb_pcls * p, gevb_field * const * fields, int nfields)
{
	gevb_ctx * c = p->ctx;
	const int b = p->cur;
	memset(&P, 0, sizeof(P));
	P.G = p->geom; P.nranks = c->nranks;
	P.pow2 = (c->N & (c->N - 1)) == 0;
	P.plane = c->plane(); P.dx = 1.0 / (double) c->N; P.rN = (double) c->N;
	P.phi = nfields >= 1 ? fields[0]->data : NULL;
	P.chi = nfields >= 2 ? fields[1]->data : NULL;
	P.B = nfields >= 3 ? fields[2]->data : NULL;
	P.csB = nfields >= 3 ? fields[2]->comp_stride : 0;
	P.nfmax = nfields;
	P.tma = gevb_tune(TUNE_GEODESIC_TMA) != 0 && nfields > 0;
	P.n = p->n;
	P.x = p->x[b]; P.y = p->y[b]; P.z = p->z[b]; P.qx = p->qx[b]; P.qy = p->qy[b]; P.qz = p->qz[b]; P.id = p->id[b]; P.key = p->key;
	P.cell_start = p->cell_start; P.cell_count = p->cell_count;
	P.rank = (gevb_tune(TUNE_REBIN_VARIANT) & 1) != 0 ? p->rank : NULL;
	P.maxv2 = (unsigned long long *) (c->d_red + 4008);
	P.nsend = (unsigned long long *) (c->d_red + 4010);
}

template <int MODE, int THREADS, int NBUF, int MINB, int PREF>
int launch_variant(gevb_pcls * p, const GParams & P)
{
	gevb_ctx * c = p->ctx;
	const int ncomp = P.nfmax >= 3 ? 5 : (P.nfmax > 0 ? P.nfmax : 1);
	const size_t smem = ((size_t) NBUF * ncomp * TILE_SITES + (PREF == 1 ? 6 * THREADS : 0)) * sizeof(double) + 16;
	GMaps M;
	memset(&M, 0, sizeof(M));
	if (P.tma && NBUF == 1)
	{
		const double * base[5] = {P.phi, P.chi, P.B, P.B + P.csB, P.B + 2 * P.csB};
		for (int f = 0; f < (P.nfmax >= 3 ? 5 : P.nfmax); f++) GEVB_TRY(gevb_tensor_map_3d(c, &M.m[f], base[f], TX, TY, TZ));
	}
	CUDA_TRY(cudaFuncSetAttribute(k_geodesic<MODE, THREADS, NBUF, MINB, PREF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	const uint32_t persistent = (uint32_t) c->num_sms * MINB;
	k_geodesic<MODE, THREADS, NBUF, MINB, PREF><<<P.G.nbricks < persistent ? P.G.nbricks : persistent, THREADS, smem, c->stream>>>(P, M);
	KERNEL_CHECK(c);
	return 0;
}

// tuning knob geodesic_variant: 0 = 256 threads, double-buffered tile, 2 blocks per SM; 1 = 128 threads, single tile buffer, 4 blocks per
// SM; 4 = the same with the next particle fetched by cp.async into shared memory (measured slower: 10.4 vs 7.7 ms at 512^3,
// profiles/r1q); 5 / 6 = the next particle in a second set of registers, 4 / 3 blocks per SM; 7 = 96 threads, 4 blocks; 2 = 256 threads,
// single buffer, 2 blocks per SM.  Default 5: with the tiles fetched by TMA the staging code no longer needs the registers that made 6
// (3 blocks, 164 registers) the better choice in round 1 -- 6.3 vs 6.7 ms per cycle (profiles/round2_v15)
template <int MODE>
int launch_geodesic(gevb_pcls * p, const GParams & P)
{
	switch (gevb_tune(TUNE_GEODESIC_VARIANT))
	{
		case 0: return launch_variant<MODE, 256, 2, 2, 0>(p, P);
		case 2: return launch_variant<MODE, 256, 1, 2, 0>(p, P);
		case 4: return launch_variant<MODE, 128, 1, 4, 1>(p, P);
		case 1: return launch_variant<MODE, 128, 1, 4, 0>(p, P);
		case 7: return launch_variant<MODE, 96, 1, 4, 2>(p, P);
		case 6: return launch_variant<MODE, 128, 1, 3, 2>(p, P);
		default: return launch_variant<MODE, 128, 1, 4, 2>(p, P);     // 5
	}
}

// what the ranks tell each other after a drift over peer memory: how many records they stored into the neighbour's slot.
// cnt (this rank's d_red area): [0,1] = my sends (down, up).  The receiver's slot ends with two counters: [0] what its
// upper neighbour sent down, [1] what its lower neighbour sent up.  The last thread also leaves the number of records the
// re-bin has to look at (live + received, clamped as k_append_received clamps) in cnt[5].
__global__ void k_publish_counts(const unsigned long long * cnt, unsigned long long * dn_counts, unsigned long long * up_counts)
{
	if (threadIdx.x == 0) dn_counts[0] = cnt[0];             // I am the upper neighbour of `dn`
	if (threadIdx.x == 1) up_counts[1] = cnt[1];             // ... and the lower neighbour of `up`
}
__global__ void k_total_records(unsigned long long * cnt, const unsigned long long * mine, unsigned long long n_live, unsigned long long cap, unsigned long long room)
{
	unsigned long long a = __ldcg(mine), b = __ldcg(mine + 1);
	a = a < cap ? a : cap; b = b < cap ? b : cap;
	cnt[2] = a; cnt[3] = b;
	unsigned long long total = n_live + a + b;
	cnt[5] = total < room ? total : room;
}

// migration over peer memory: the drift kernel stored the leavers straight into the neighbours' slots; publish the
// counts, barrier, append what arrived (counts stay on the device), re-file, then ONE read-back for the host
int finish_move_peer(gevb_pcls * p, GParams & P, int slot)
{
	gevb_ctx * c = p->ctx;
	Timed timed_mig_(c, CLS_MIGRATE);
	const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
	const int64_t cap = (int64_t) c->pc_mig_cap;
	unsigned long long * cnt = P.nsend;
	auto counters = [&](int r) { return (unsigned long long *) (gevb_peer_slot(c, r, slot) + c->pc_plane_doubles + 2 * 7 * c->pc_mig_cap); };
	k_publish_counts<<<1, 32, 0, c->stream>>>(cnt, counters(dn), counters(up));
	KERNEL_CHECK(c);
	GEVB_TRY(gevb_peer_barrier(c, c->stream));
	const int b = p->cur;
	const double dx = 1.0 / (double) c->N;
	const double * mig = gevb_peer_slot(c, c->rank, slot) + c->pc_plane_doubles;
	k_total_records<<<1, 1, 0, c->stream>>>(cnt, counters(c->rank), (unsigned long long) p->n, (unsigned long long) cap, (unsigned long long) p->cap);
	KERNEL_CHECK(c);
	const int grid = gevb_grid(c, (size_t) (cap < p->cap - p->n ? cap : (p->cap - p->n > 0 ? p->cap - p->n : 1)), 256, 4);
	for (int which = 0; which < 2; which++)
	{
		k_append_received<<<grid, 256, 0, c->stream>>>(0, mig + (size_t) which * 7 * cap, cap, p->n, p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], p->key, p->geom, dx,
			p->cell_count, cnt + 4, P.rank ? p->rank : NULL, counters(c->rank), which, p->cap);
		KERNEL_CHECK(c);
	}
	// the departed carry GEVB_INVALID_KEY and are dropped by the move; it reads the number of records from the device
	p->d_nin = cnt + 5;
	const int64_t n_upper = p->n + 2 * cap < p->cap ? p->n + 2 * cap : p->cap;
	const int r = gevb_pcls_rebin(p, n_upper, p->n, true);
	p->d_nin = NULL;
	GEVB_TRY(r);
	// one read-back: sends, receives, lost, overflow
	CUDA_TRY(cudaMemcpyAsync(c->h_red + 16, cnt, 7 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	const unsigned long long * h = (const unsigned long long *) (c->h_red + 16);
	GEVB_TRY(gevb_peer_check(c));
	GEVB_CHECK_ARG((int64_t) h[0] <= cap && (int64_t) h[1] <= cap, "moveParticles: migration buffer overflow (%llu / %llu sent, capacity %lld per direction; set GEVB_MIGRATION_CAP)", h[0], h[1], (long long) cap);
	GEVB_CHECK_ARG(h[4] == 0, "moveParticles: %llu particles moved farther than the adjacent slab (move limit, main.cpp:281-286)", h[4]);
	GEVB_CHECK_ARG(h[6] == 0 && p->n + (int64_t) (h[2] + h[3]) <= p->cap, "moveParticles: received particles do not fit the particle arrays (%lld + %llu > %lld)", (long long) p->n, h[2] + h[3], (long long) p->cap);
	p->n = p->n + (int64_t) (h[2] + h[3]) - (int64_t) (h[0] + h[1]);
	return 0;
}

// after a drift: exchange slab-crossing particles with the ring neighbours, then re-file (counting sort)
int finish_move(gevb_pcls * p, GParams & P, int peer_slot)
{
	gevb_ctx * c = p->ctx;
	if (c->nranks == 1) { Timed timed_(c, CLS_SORT); return gevb_pcls_rebin(p, p->n, p->n, true); }
	if (peer_slot >= 0) return finish_move_peer(p, P, peer_slot);
	Timed timed_mig_(c, CLS_MIGRATE);
	const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
	unsigned long long * cnt = P.nsend;                 // [0,1] = my sends (down, up); [2,3] = what I receive (from up, from down)
	NCCL_TRY(ncclGroupStart());
	NCCL_TRY(ncclSend(cnt + 0, 1, ncclUint64, dn, c->comm, c->stream));
	NCCL_TRY(ncclSend(cnt + 1, 1, ncclUint64, up, c->comm, c->stream));
	NCCL_TRY(ncclRecv(cnt + 2, 1, ncclUint64, up, c->comm, c->stream));     // what my upper neighbour sends down to me
	NCCL_TRY(ncclRecv(cnt + 3, 1, ncclUint64, dn, c->comm, c->stream));     // what my lower neighbour sends up to me
	NCCL_TRY(ncclGroupEnd());
	unsigned long long h[4];
	CUDA_TRY(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	GEVB_CHECK_ARG((int64_t) h[0] <= P.sendcap && (int64_t) h[1] <= P.sendcap && (int64_t) h[2] <= P.sendcap && (int64_t) h[3] <= P.sendcap,
		"moveParticles: migration buffer overflow (%llu/%llu sent, %llu/%llu received, capacity %lld)", h[0], h[1], h[2], h[3], (long long) P.sendcap);
	const int64_t nrecv = (int64_t) (h[2] + h[3]);
	// receive staging lives behind the two send buffers
	double * rb_up = P.sendbuf[1] + 7 * P.sendcap, * rb_dn = rb_up + 7 * P.sendcap;
	NCCL_TRY(ncclGroupStart());
	for (int a = 0; a < 7; a++)
	{
		if (h[0]) NCCL_TRY(ncclSend(P.sendbuf[0] + a * P.sendcap, h[0], ncclDouble, dn, c->comm, c->stream));
		if (h[1]) NCCL_TRY(ncclSend(P.sendbuf[1] + a * P.sendcap, h[1], ncclDouble, up, c->comm, c->stream));
		if (h[2]) NCCL_TRY(ncclRecv(rb_up + a * P.sendcap, h[2], ncclDouble, up, c->comm, c->stream));
		if (h[3]) NCCL_TRY(ncclRecv(rb_dn + a * P.sendcap, h[3], ncclDouble, dn, c->comm, c->stream));
	}
	NCCL_TRY(ncclGroupEnd());
	if (p->n + nrecv > p->cap)
	{
		// reserve() copies the live arrays (keys included) and resets cur to 0
		GEVB_TRY(gevb_pcls_reserve(p, p->n + nrecv + (p->n + nrecv) / 16));
	}
	const int b = p->cur;
	const double dx = 1.0 / (double) c->N;
	if (h[2])
	{
		k_append_received<<<gevb_grid(c, h[2], 256), 256, 0, c->stream>>>((int64_t) h[2], rb_up, P.sendcap, p->n, p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], p->key, p->geom, dx, p->cell_count, cnt + 4, P.rank ? p->rank : NULL);
		KERNEL_CHECK(c);
	}
	if (h[3])
	{
		k_append_received<<<gevb_grid(c, h[3], 256), 256, 0, c->stream>>>((int64_t) h[3], rb_dn, P.sendcap, p->n + (int64_t) h[2], p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], p->key, p->geom, dx, p->cell_count, cnt + 4, P.rank ? p->rank : NULL);
		KERNEL_CHECK(c);
	}
	// the departed carry GEVB_INVALID_KEY and are dropped by the scatter
	GEVB_TRY(gevb_pcls_rebin(p, p->n + nrecv, p->n + nrecv - (int64_t) (h[0] + h[1]), true));
	unsigned long long lost = 0;
	CUDA_TRY(cudaMemcpyAsync(&lost, cnt + 4, sizeof(lost), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	GEVB_CHECK_ARG(lost == 0, "moveParticles: %llu particles moved farther than the adjacent slab (move limit, main.cpp:281-286)", lost);
	return 0;
}

// returns in *peer_slot the slot of the peer buffers the leavers are stored into, or -1 (NCCL messages from local buffers)
int setup_migration(gevb_pcls * p, GParams & P, int * peer_slot)
{
	gevb_ctx * c = p->ctx;
	*peer_slot = -1;
	if (c->nranks == 1) return 0;
	if (gevb_peer_on(c))
	{
		// the drift kernel stores straight into the neighbours' memory: moving down -> region 0 of the lower neighbour's slot
		// ("what my upper neighbour sent down"), moving up -> region 1 of the upper neighbour's slot
		const int up = (c->rank + 1) % c->nranks, dn = (c->rank + c->nranks - 1) % c->nranks;
		const int slot = (int) (c->pc_seq++ & 1);
		P.sendcap = (int64_t) c->pc_mig_cap;
		P.sendbuf[0] = gevb_peer_slot(c, dn, slot) + c->pc_plane_doubles;
		P.sendbuf[1] = gevb_peer_slot(c, up, slot) + c->pc_plane_doubles + 7 * c->pc_mig_cap;
		// room for what may arrive (the counts stay on the device until the end of the call)
		// (grown with hysteresis: the slab population drifts by a few particles per step, which must not reallocate every step)
		const int64_t need = p->n + p->n / 8 + 65536, want = p->n + p->n / 4 + 131072;
		if (p->cap < need) { GEVB_TRY(gevb_pcls_reserve(p, want)); P.x = p->x[p->cur]; P.y = p->y[p->cur]; P.z = p->z[p->cur]; P.qx = p->qx[p->cur]; P.qy = p->qy[p->cur]; P.qz = p->qz[p->cur]; P.id = p->id[p->cur]; P.key = p->key; if (P.rank) P.rank = p->rank; }
		CUDA_TRY(cudaMemsetAsync(P.nsend, 0, 7 * sizeof(unsigned long long), c->stream));
		*peer_slot = slot;
		return 0;
	}
	int64_t cap = p->n / 8 + 65536;
	void * buf;
	GEVB_TRY(gevb_ctx_scratch2(c, (size_t) cap * 7 * 4 * sizeof(double), &buf));
	P.sendcap = cap;
	P.sendbuf[0] = (double *) buf;
	P.sendbuf[1] = (double *) buf + 7 * cap;
	CUDA_TRY(cudaMemsetAsync(P.nsend, 0, 5 * sizeof(unsigned long long), c->stream));
	return 0;
}

} // namespace

extern "C" int gevb_updateVel(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params, double * maxvel)
{
	static const double no_params[2] = {1., 1.};
	GEVB_CHECK_ARG(p != NULL, "updateVel: NULL argument");
	GEVB_CHECK_ARG(fn == GEVB_UPDATE_Q || fn == GEVB_UPDATE_Q_NEWTON || fn == GEVB_INITIALIZE_Q_IC_BASIC, "updateVel: unknown callback %d (only update_q, update_q_Newton and initialize_q_ic_basic can run on the device)", fn);
	if (fn == GEVB_INITIALIZE_Q_IC_BASIC) { params = no_params; GEVB_CHECK_ARG(nfields <= 2, "updateVel: initialize_q_ic_basic takes one or two potentials"); }
	GEVB_CHECK_ARG(params != NULL, "updateVel: NULL params");
	gevb_ctx * c = p->ctx;
	GEVB_TRY(check_fields(c, fields, nfields, "updateVel"));
	GEVB_CHECK_ARG(nfields >= 1, "updateVel: needs at least phi");
	CUDA_TRY(cudaSetDevice(c->device));
	GParams P;
	base_params(P, p, fields, nfields);
	P.fn = fn; P.nf_kick = nfields; P.dtau_kick = dtau; P.a_kick = params[0]; P.bscale_kick = params[1]; P.binv_kick = 1.0 / params[1];
	CUDA_TRY(cudaMemsetAsync(P.maxv2, 0, sizeof(unsigned long long), c->stream));
	if (p->n > 0)
	{
		Timed timed_(c, CLS_KICK);
		GEVB_TRY(launch_geodesic<0>(p, P));
	}
	if (maxvel)
	{
		CUDA_TRY(cudaMemcpyAsync(c->h_red, P.maxv2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*maxvel = sqrt(c->h_red[0] / params[0] / params[0]);   // callback returns v2/a/a (gevolution.hpp:670); LATfield2 updateVel returns sqrt(max), cf. main.cpp:816-822
	}
	return 0;
}

static int move_particles(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params, double * output_max);

extern "C" int gevb_moveParticles(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params)
{
	return move_particles(p, fn, dtau, fields, nfields, params, NULL);
}

extern "C" int gevb_moveParticles_max(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params, double * output_max)
{
	GEVB_CHECK_ARG(output_max != NULL, "moveParticles: NULL output");
	return move_particles(p, fn, dtau, fields, nfields, params, output_max);
}

static int move_particles(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params, double * output_max)
{
	static const double no_params[2] = {1., 1.};
	GEVB_CHECK_ARG(p != NULL, "moveParticles: NULL argument");
	GEVB_CHECK_ARG(fn == GEVB_UPDATE_Q || fn == GEVB_UPDATE_Q_NEWTON || fn == GEVB_DISPLACE_PCLS_IC_BASIC, "moveParticles: unknown callback %d (only update_pos, update_pos_Newton and displace_pcls_ic_basic can run on the device)", fn);
	if (fn == GEVB_DISPLACE_PCLS_IC_BASIC) { params = no_params; GEVB_CHECK_ARG(nfields >= 1 && nfields <= 2, "moveParticles: displace_pcls_ic_basic takes one or two displacement fields"); }
	GEVB_CHECK_ARG(params != NULL, "moveParticles: NULL params");
	GEVB_CHECK_ARG(output_max == NULL || fn == GEVB_DISPLACE_PCLS_IC_BASIC, "moveParticles: only displace_pcls_ic_basic has a reduction output");
	gevb_ctx * c = p->ctx;
	if (fn == GEVB_UPDATE_Q_NEWTON) nfields = 0;
	GEVB_TRY(check_fields(c, fields, nfields, "moveParticles"));
	CUDA_TRY(cudaSetDevice(c->device));
	GParams P;
	base_params(P, p, fields, nfields);
	P.fn = fn; P.nf_drift = nfields; P.dtau_drift = dtau; P.a_drift = params[0]; P.bscale_drift = params[1]; P.binv_drift = 1.0 / params[1];
	int peer_slot = -1;
	GEVB_TRY(setup_migration(p, P, &peer_slot));
	CUDA_TRY(cudaMemsetAsync(P.maxv2, 0, sizeof(unsigned long long), c->stream));
	if (p->n > 0)
	{
		Timed timed_(c, CLS_DRIFT);
		GEVB_TRY(launch_geodesic<1>(p, P));
	}
	GEVB_TRY(finish_move(p, P, peer_slot));
	if (output_max)
	{
		// the callback's reduction output (largest displacement, ic_basic.hpp:88-89) of the LOCAL particles; the caller reduces over ranks
		CUDA_TRY(cudaMemcpyAsync(c->h_red, P.maxv2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*output_max = sqrt(c->h_red[0]);
	}
	return 0;
}

extern "C" int gevb_kick_drift(gevb_pcls * p, int fn, double dtau_kick, int nfields_kick, const double * params_kick,
                               double dtau_drift, int nfields_drift, const double * params_drift,
                               gevb_field * const * fields, double * maxvel)
{
	GEVB_CHECK_ARG(p != NULL && params_kick != NULL && params_drift != NULL, "kick_drift: NULL argument");
	GEVB_CHECK_ARG(fn == GEVB_UPDATE_Q || fn == GEVB_UPDATE_Q_NEWTON, "kick_drift: unknown callback %d", fn);
	gevb_ctx * c = p->ctx;
	if (fn == GEVB_UPDATE_Q_NEWTON) nfields_drift = 0;
	const int nf = nfields_kick > nfields_drift ? nfields_kick : nfields_drift;
	GEVB_TRY(check_fields(c, fields, nf, "kick_drift"));
	GEVB_CHECK_ARG(nfields_kick >= 1, "kick_drift: the kick needs at least phi");
	CUDA_TRY(cudaSetDevice(c->device));
	GParams P;
	base_params(P, p, fields, nf);
	P.fn = fn;
	P.nf_kick = nfields_kick; P.dtau_kick = dtau_kick; P.a_kick = params_kick[0]; P.bscale_kick = params_kick[1]; P.binv_kick = 1.0 / params_kick[1];
	P.nf_drift = nfields_drift; P.dtau_drift = dtau_drift; P.a_drift = params_drift[0]; P.bscale_drift = params_drift[1]; P.binv_drift = 1.0 / params_drift[1];
	int peer_slot = -1;
	GEVB_TRY(setup_migration(p, P, &peer_slot));
	CUDA_TRY(cudaMemsetAsync(P.maxv2, 0, sizeof(unsigned long long), c->stream));
	if (p->n > 0)
	{
		Timed timed_(c, CLS_KICK_DRIFT);
		GEVB_TRY(launch_geodesic<2>(p, P));
	}
	// the max lives in its own reduction slot, untouched by finish_move
	GEVB_TRY(finish_move(p, P, peer_slot));
	if (maxvel)
	{
		CUDA_TRY(cudaMemcpyAsync(c->h_red, P.maxv2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*maxvel = sqrt(c->h_red[0] / params_kick[0] / params_kick[0]);
	}
	return 0;
}
