// deposit.cu -- particle -> mesh projections (CIC / staggered CIC-NGP)
//
//   projection_T00_project       gevolution.hpp:927-1022
//   projection_T0i_project       gevolution.hpp:1046-1147
//   projection_Tij_project       gevolution.hpp:1173-1297
//   scalarProjectionCIC_project  LATfield2 (main.cpp:402), plain CIC
//
// One thread block per brick of 16 x 8 x 4 cells (particles are stored brick by
// brick and, inside a brick, cell by cell -- gevb_internal.cuh).  The block owns a
// 17 x 9 x 5-site shared-memory tile per target component (the brick's sites plus
// the upper apron the CIC cloud reaches) and a tile of phi of the same shape.
//
//   1. every thread owns cells of the brick and accumulates the contributions of
//      the cell's particles in registers -- exactly the reference's per-cell
//      "localCube" accumulators (gevolution.hpp:953,1199-1200); particle loads are
//      coalesced because consecutive threads own consecutive cells = consecutive
//      particles.  Cells holding more than DEP_LIGHT particles hand the excess to
//      a warp-cooperative pass (strided loads, shuffle reduction), so clustered
//      states do not serialise on one thread.
//   2. the per-cell sums go into the tile in eight corner phases: in one phase
//      all threads write the same corner of their own cell, i.e. distinct sites,
//      so the shared-memory update is a plain read-add-write, no atomics.
//   3. the tile is flushed to the field in HBM with FP64 reductions (RED.ADD.F64),
//      one per non-zero tile site and component -- about 5 k per brick instead of
//      38 per particle; consecutive lanes flush consecutive sites of a row.
//
// x and y wrap by index arithmetic at flush time; z+1 of the last local plane
// lands in the upper ghost plane which gevb_projection_comm folds into the next rank.
#include "gevb_internal.cuh"

namespace {

#define DX (GEVB_BX + 1)
#define DY (GEVB_BY + 1)
#define DZ (GEVB_BZ + 1)
#define DT_SITES (DX * DY * DZ)
#define DEP_LIGHT 4
#define DEP_THREADS 256

enum { DEP_T00 = 0, DEP_TIJ = 1, DEP_T00_TIJ = 2, DEP_T0I = 3 };

// accumulators per cell and tile components per projection
__host__ __device__ constexpr int dep_nacc(int what) { return what == DEP_T00 ? 8 : what == DEP_TIJ ? 30 : what == DEP_T00_TIJ ? 38 : 12; }
__host__ __device__ constexpr int dep_ncomp(int what) { return what == DEP_T00 ? 1 : what == DEP_TIJ ? 6 : what == DEP_T00_TIJ ? 7 : 3; }

// symmetric-tensor accumulators: 24 diagonal (tii[k + 8 d], gevolution.hpp:1199) then 6 off-diagonal (tij[0..5], :1198)
__host__ __device__ constexpr int tij_comp(int a) { return a < 24 ? (a / 8 == 0 ? 0 : a / 8 == 1 ? 3 : 5) : (a - 24 < 2 ? 1 : a - 24 < 4 ? 2 : 4); }
__host__ __device__ constexpr int tij_corner(int a) { return a < 24 ? a % 8 : (a == 24 || a == 26 || a == 28 ? 0 : a == 25 ? 1 : a == 27 ? 2 : 4); }   // :1277-1294
// T0i accumulators qi[0..11] (gevolution.hpp:1107-1126): component a / 4, corner from the write-back at :1129-1144
__host__ __device__ constexpr int t0i_corner(int a)
{
	return a == 0 || a == 4 || a == 8 ? 0 : a == 1 ? 2 : a == 2 ? 1 : a == 3 ? 3 : a == 5 ? 4 : a == 6 ? 1 : a == 7 ? 5 : a == 9 ? 4 : a == 10 ? 2 : 6;
}
// tile component and corner (4X + 2Y + Z, gevolution.hpp:953) of accumulator a
__host__ __device__ constexpr int acc_comp(int what, int a)
{
	return what == DEP_T00 ? 0 : what == DEP_TIJ ? tij_comp(a) : what == DEP_T00_TIJ ? (a < 8 ? 0 : 1 + tij_comp(a - 8)) : a / 4;
}
__host__ __device__ constexpr int acc_corner(int what, int a)
{
	return what == DEP_T00 ? a : what == DEP_TIJ ? tij_corner(a) : what == DEP_T00_TIJ ? (a < 8 ? a : tij_corner(a - 8)) : t0i_corner(a);
}
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

// contributions of one particle to the cell's accumulators
template <int WHAT, bool HAS_PHI>
__device__ __forceinline__ void accumulate(double * acc, const DParams & D, const double * pv, double refx, double refy, double refz, const double * cphi)
{
	double up[3], dn[3];
	if (D.pow2)
	{
		up[0] = (pv[0] - refx) * D.rN; up[1] = (pv[1] - refy) * D.rN; up[2] = (pv[2] - refz) * D.rN;      // == / dx exactly (dx = 2^-k)
	}
	else
	{
		up[0] = (pv[0] - refx) / D.dx; up[1] = (pv[1] - refy) / D.dx; up[2] = (pv[2] - refz) / D.dx;      // gevolution.hpp:981 / :1101 / :1231
	}
	dn[0] = 1.0 - up[0]; dn[1] = 1.0 - up[1]; dn[2] = 1.0 - up[2];                                          // :982
	const double q0 = pv[3], q1 = pv[4], q2 = pv[5];
	if (WHAT == DEP_T0I)
	{
		double w = D.mass * q0;                                                    // :1107
		acc[0] += w * dn[1] * dn[2]; acc[1] += w * up[1] * dn[2]; acc[2] += w * dn[1] * up[2]; acc[3] += w * up[1] * up[2];       // :1109-1112
		w = D.mass * q1;                                                           // :1114
		acc[4] += w * dn[0] * dn[2]; acc[5] += w * up[0] * dn[2]; acc[6] += w * dn[0] * up[2]; acc[7] += w * up[0] * up[2];       // :1116-1119
		w = D.mass * q2;                                                           // :1121
		acc[8] += w * dn[0] * dn[1]; acc[9] += w * up[0] * dn[1]; acc[10] += w * dn[0] * up[1]; acc[11] += w * up[0] * up[1];     // :1123-1126
		return;
	}
	const double qsq = q0 * q0 + q1 * q1 + q2 * q2;
	double w[8];
	#pragma unroll
	for (int k = 0; k < 8; k++) w[k] = ((k & 4) ? up[0] : dn[0]) * ((k & 2) ? up[1] : dn[1]) * ((k & 1) ? up[2] : dn[2]);
	if (WHAT == DEP_T00 || WHAT == DEP_T00_TIJ)
	{
		double e = D.a, f = 0.;
		if (HAS_PHI) { e = sqrt(qsq + D.a * D.a); f = 3. * e + qsq / e; }          // :989-991
		#pragma unroll
		for (int k = 0; k < 8; k++) acc[k] += w[k] * (e + f * cphi[k]);            // :995-1009 (mass applied at write-back, :1012)
	}
	if (WHAT == DEP_TIJ || WHAT == DEP_T00_TIJ)
	{
		double * t = acc + (WHAT == DEP_T00_TIJ ? 8 : 0);
		const double e = sqrt(qsq + D.a * D.a);                                    // :1237
		const double f = 4. + D.a * D.a / (qsq + D.a * D.a);                       // :1238
		const double qq[3] = {q0, q1, q2};
		double g[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) g[k] = w[k] * (1. + f * cphi[k]);
		#pragma unroll
		for (int d = 0; d < 3; d++)
		{
			const double wd = D.mass * qq[d] * qq[d] / e;                          // :1243
			#pragma unroll
			for (int k = 0; k < 8; k++) t[d * 8 + k] += wd * g[k];                 // :1245-1259
		}
		double wo = D.mass * q0 * q1 / e;                                          // :1262
		t[24] += wo * dn[2] * (1. + f * 0.25 * (cphi[0] + cphi[2] + cphi[4] + cphi[6]));
		t[25] += wo * up[2] * (1. + f * 0.25 * (cphi[1] + cphi[3] + cphi[5] + cphi[7]));
		wo = D.mass * q0 * q2 / e;                                                 // :1266
		t[26] += wo * dn[1] * (1. + f * 0.25 * (cphi[0] + cphi[1] + cphi[4] + cphi[5]));
		t[27] += wo * up[1] * (1. + f * 0.25 * (cphi[2] + cphi[3] + cphi[6] + cphi[7]));
		wo = D.mass * q1 * q2 / e;                                                 // :1270
		t[28] += wo * dn[0] * (1. + f * 0.25 * (cphi[0] + cphi[1] + cphi[2] + cphi[3]));
		t[29] += wo * up[0] * (1. + f * 0.25 * (cphi[4] + cphi[5] + cphi[6] + cphi[7]));
	}
}

// per-cell factors the reference applies at write-back
template <int WHAT>
__device__ __forceinline__ void finalize(double * acc, const DParams & D, const double * cphi)
{
	if (WHAT == DEP_T00 || WHAT == DEP_T00_TIJ)
	{
		#pragma unroll
		for (int k = 0; k < 8; k++) acc[k] *= D.mass;                              // :1012-1019
	}
	if (WHAT == DEP_T0I)
	{
		#pragma unroll
		for (int a = 0; a < 12; a++)
		{
			const int k = t0i_corner(a), own = a / 4 == 0 ? 4 : a / 4 == 1 ? 2 : 1;   // the component's own axis bit
			acc[a] *= 1. + cphi[k] + cphi[k + own];                                // :1129-1144
		}
	}
}

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

#define DEP_CELLTAB (GEVB_BRICK_CELLS + 4)      // 513 prefix sums of a brick, padded
#define DEP_PCAP 512                            // particles of a brick staged in shared memory (the rest is read from HBM directly)
#define DEP_STAGE_DOUBLES (DT_SITES + 6 * DEP_PCAP + DEP_CELLTAB / 2)   // one pipeline stage: phi tile, 6 particle arrays, cell table

// asynchronous copy (LDGSTS) of everything a brick's deposit reads: its slice of cell_start[], the phi tile and the
// first DEP_PCAP particles.  For the tile a thread keeps its (tx, ty) column and walks z.
template <bool HAS_PHI>
__device__ __forceinline__ void stage_brick(const DParams & D, uint32_t brick, uint32_t first, uint32_t last, double * stage)
{
	const BrickGeom & G = D.G;
	double * tphi = stage, * part = stage + DT_SITES;
	uint32_t * ctab = (uint32_t *) (stage + DT_SITES + 6 * DEP_PCAP);
	if (first == last) return;
	const uint32_t * src = D.cell_start + (size_t) brick * GEVB_BRICK_CELLS;
	for (int k = threadIdx.x; k <= GEVB_BRICK_CELLS; k += DEP_THREADS) cp_async4(ctab + k, src + k);
	const uint32_t n = last - first < DEP_PCAP ? last - first : DEP_PCAP;
	for (uint32_t k = threadIdx.x; k < n; k += DEP_THREADS)
	{
		cp_async8(part + k, D.x + first + k); cp_async8(part + DEP_PCAP + k, D.y + first + k); cp_async8(part + 2 * DEP_PCAP + k, D.z + first + k);
		cp_async8(part + 3 * DEP_PCAP + k, D.qx + first + k); cp_async8(part + 4 * DEP_PCAP + k, D.qy + first + k); cp_async8(part + 5 * DEP_PCAP + k, D.qz + first + k);
	}
	if (HAS_PHI && threadIdx.x < DX * DY)
	{
		int x0, y0, zl0;
		brick_origin(G, brick, x0, y0, zl0);
		const int tx = threadIdx.x % DX, ty = threadIdx.x / DX;
		const size_t gcol = (size_t) ((y0 + ty) % G.N) * G.N + (x0 + tx) % G.N;
		#pragma unroll
		for (int tz = 0; tz < DZ; tz++)
		{
			const int plane = zl0 + tz + 1;
			if (plane > G.nzl + 1) break;                       // partial brick at the top of the slab: never read
			cp_async8(tphi + (tz * DY + ty) * DX + tx, D.phi + (size_t) plane * G.N * G.N + gcol);
		}
	}
}

// particle i of the current brick: from the staged copy if it is there, else from HBM
__device__ __forceinline__ void load_particle(const DParams & D, const double * part, uint32_t bfirst, uint32_t i, double * pv)
{
	const uint32_t k = i - bfirst;
	if (k < DEP_PCAP)
	{
		#pragma unroll
		for (int a = 0; a < 6; a++) pv[a] = part[a * DEP_PCAP + k];
	}
	else
	{
		pv[0] = D.x[i]; pv[1] = D.y[i]; pv[2] = D.z[i]; pv[3] = D.qx[i]; pv[4] = D.qy[i]; pv[5] = D.qz[i];
	}
}

__device__ __forceinline__ void brick_range(const DParams & D, uint32_t b, uint32_t & first, uint32_t & last)
{
	first = last = 0;
	if (b < D.G.nbricks) { first = __ldg(D.cell_start + (size_t) b * GEVB_BRICK_CELLS); last = __ldg(D.cell_start + (size_t) (b + 1) * GEVB_BRICK_CELLS); }
}

// Persistent blocks walk the bricks with stride gridDim.x in a two-stage software pipeline: while brick k is
// accumulated and flushed, everything brick k+1 needs is in flight into the other shared-memory stage, and the
// particle range of brick k+2 is being fetched into registers.
template <int WHAT, bool HAS_PHI>
__global__ void __launch_bounds__(DEP_THREADS, 2) k_deposit(DParams D)
{
	constexpr int NACC = dep_nacc(WHAT), NCOMP = dep_ncomp(WHAT);
	extern __shared__ double smem[];
	double * tile = smem;                                   // [NCOMP][DT_SITES] accumulators
	double * stages = smem + NCOMP * DT_SITES;              // [2][DEP_STAGE_DOUBLES]
	__shared__ uint16_t heavy_cell[GEVB_BRICK_CELLS];
	__shared__ int nheavy;

	const BrickGeom & G = D.G;
	const int tcol = threadIdx.x, ttx = tcol % DX, tty = tcol / DX;      // this thread's column of the tile (flush)
	for (int idx = threadIdx.x; idx < NCOMP * DT_SITES + 2 * DEP_STAGE_DOUBLES; idx += DEP_THREADS) smem[idx] = 0.;
	if (threadIdx.x == 0) nheavy = 0;
	__syncthreads();
	uint32_t brick = blockIdx.x;
	uint32_t first, last, nfirst, nlast, nnfirst, nnlast;
	brick_range(D, brick, first, last);
	brick_range(D, brick + gridDim.x, nfirst, nlast);
	stage_brick<HAS_PHI>(D, brick, first, last, stages);
	cp_async_commit();
	int cur = 0;
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
			const double * part = tphi + DT_SITES;
			const uint32_t * ctab = (const uint32_t *) (part + 6 * DEP_PCAP);

			// ---- light pass: one thread per cell, DEP_LIGHT particles at most -------------------------
			#pragma unroll 1
			for (int cc = 0; cc < GEVB_BRICK_CELLS / DEP_THREADS; cc++)
			{
				const int c = cc * DEP_THREADS + threadIdx.x;
				const int sx = c & (GEVB_BX - 1), sy = (c >> GEVB_BX_BITS) & (GEVB_BY - 1), sz = c >> (GEVB_BX_BITS + GEVB_BY_BITS);
				const int site = (sz * DY + sy) * DX + sx;
				const uint32_t cfirst = ctab[c], clast = ctab[c + 1];
				const uint32_t n = clast - cfirst;
				double acc[NACC];
				#pragma unroll
				for (int a = 0; a < NACC; a++) acc[a] = 0.;
				if (n > 0)
				{
					double cphi[8];
					#pragma unroll
					for (int k = 0; k < 8; k++) cphi[k] = tphi[site + corner_offset(k)];   // :967-974
					const double refx = (x0 + sx) * D.dx, refy = (y0 + sy) * D.dx, refz = (G.z0 + zl0 + sz) * D.dx;   // referPos, :963
					const uint32_t nl = n < DEP_LIGHT ? n : DEP_LIGHT;
					for (uint32_t j = 0; j < nl; j++)
					{
						double pv[6];
						load_particle(D, part, first, cfirst + j, pv);
						accumulate<WHAT, HAS_PHI>(acc, D, pv, refx, refy, refz, cphi);
					}
					finalize<WHAT>(acc, D, cphi);
					if (n > DEP_LIGHT)
					{
						const int h = atomicAdd(&nheavy, 1);
						heavy_cell[h] = (uint16_t) c;
					}
				}
				// eight corner phases: within a phase every thread updates a different site
				#pragma unroll
				for (int k = 0; k < 8; k++)
				{
					if (n > 0)
					{
						#pragma unroll
						for (int a = 0; a < NACC; a++)
							if (acc_corner(WHAT, a) == k) tile[acc_comp(WHAT, a) * DT_SITES + site + corner_offset(k)] += acc[a];
					}
					__syncthreads();
				}
			}

			// ---- heavy pass: one warp per crowded cell -------------------------------------------------
			const int nh = nheavy;
			for (int h = threadIdx.x >> 5; h < nh; h += DEP_THREADS / 32)
			{
				const int c = heavy_cell[h];
				const int sx = c & (GEVB_BX - 1), sy = (c >> GEVB_BX_BITS) & (GEVB_BY - 1), sz = c >> (GEVB_BX_BITS + GEVB_BY_BITS);
				const int site = (sz * DY + sy) * DX + sx;
				double acc[NACC];
				#pragma unroll
				for (int a = 0; a < NACC; a++) acc[a] = 0.;
				double cphi[8];
				#pragma unroll
				for (int k = 0; k < 8; k++) cphi[k] = tphi[site + corner_offset(k)];
				const double refx = (x0 + sx) * D.dx, refy = (y0 + sy) * D.dx, refz = (G.z0 + zl0 + sz) * D.dx;
				const uint32_t hlast = ctab[c + 1];
				for (uint32_t i = ctab[c] + DEP_LIGHT + (threadIdx.x & 31); i < hlast; i += 32)
				{
					double pv[6];
					load_particle(D, part, first, i, pv);
					accumulate<WHAT, HAS_PHI>(acc, D, pv, refx, refy, refz, cphi);
				}
				#pragma unroll
				for (int a = 0; a < NACC; a++)
					for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
				finalize<WHAT>(acc, D, cphi);
				if ((threadIdx.x & 31) == 0)
				{
					#pragma unroll
					for (int a = 0; a < NACC; a++) atomicAdd(&tile[acc_comp(WHAT, a) * DT_SITES + site + corner_offset(acc_corner(WHAT, a))], acc[a]);
				}
			}
			__syncthreads();                                // stage `cur` is free from here on
			if (threadIdx.x == 0) nheavy = 0;

			// ---- flush the tile: FP64 reductions into HBM, consecutive lanes on consecutive sites of a row;
			//      the tile is left zeroed for the next brick
			if (tcol < DX * DY)
			{
				const size_t gcol = (size_t) ((y0 + tty) % G.N) * G.N + (x0 + ttx) % G.N;
				#pragma unroll
				for (int tz = 0; tz < DZ; tz++)
				{
					const int s = (tz * DY + tty) * DX + ttx;
					const size_t off = (size_t) (zl0 + tz + 1) * G.N * G.N + gcol;
					#pragma unroll
					for (int k = 0; k < NCOMP; k++)
					{
						const double v = tile[k * DT_SITES + s];
						if (v != 0.) { atomicAdd(D.out[k] + off, v); tile[k * DT_SITES + s] = 0.; }
					}
				}
			}
		}
		brick = nbrick; cur ^= 1;
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
	const size_t smem = ((size_t) dep_ncomp(WHAT) * DT_SITES + 2 * DEP_STAGE_DOUBLES) * sizeof(double);
	const uint32_t persistent = (uint32_t) c->num_sms * 2;
	const uint32_t grid = D.G.nbricks < persistent ? D.G.nbricks : persistent;
	if (phi)
	{
		CUDA_TRY(cudaFuncSetAttribute(k_deposit<WHAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_deposit<WHAT, true><<<grid, DEP_THREADS, smem, c->stream>>>(D);
	}
	else
	{
		CUDA_TRY(cudaFuncSetAttribute(k_deposit<WHAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_deposit<WHAT, false><<<grid, DEP_THREADS, smem, c->stream>>>(D);
	}
	KERNEL_CHECK(c);
	return 0;
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
