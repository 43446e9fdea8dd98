// deposit.cu -- particle -> mesh projections (CIC / staggered CIC-NGP)
//
//   projection_T00_project       gevolution.hpp:927-1022
//   projection_T0i_project       gevolution.hpp:1046-1147
//   projection_Tij_project       gevolution.hpp:1173-1297
//   scalarProjectionCIC_project  LATfield2 (main.cpp:402), plain CIC
//
// One thread block per brick of 8^3 cells (particles are stored brick by brick
// and, inside a brick, cell by cell -- gevb_internal.cuh).  The block owns a
// 9^3-site shared-memory tile per target component (the brick's sites plus the
// upper apron the CIC cloud reaches) and a 9^3 tile of phi.
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

#define DT_EDGE 9
#define DT_SITES 729
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
__host__ __device__ constexpr int corner_offset(int k) { return ((k >> 2) & 1) + ((k >> 1) & 1) * DT_EDGE + (k & 1) * DT_EDGE * DT_EDGE; }

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
__device__ __forceinline__ void accumulate(double * acc, const DParams & D, uint32_t i, double refx, double refy, double refz, const double * cphi)
{
	double up[3], dn[3];
	if (D.pow2)
	{
		up[0] = (D.x[i] - refx) * D.rN; up[1] = (D.y[i] - refy) * D.rN; up[2] = (D.z[i] - refz) * D.rN;   // == / dx exactly (dx = 2^-k)
	}
	else
	{
		up[0] = (D.x[i] - refx) / D.dx; up[1] = (D.y[i] - refy) / D.dx; up[2] = (D.z[i] - refz) / D.dx;   // gevolution.hpp:981 / :1101 / :1231
	}
	dn[0] = 1.0 - up[0]; dn[1] = 1.0 - up[1]; dn[2] = 1.0 - up[2];                                          // :982
	const double q0 = D.qx[i], q1 = D.qy[i], q2 = D.qz[i];
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

template <int WHAT, bool HAS_PHI>
__global__ void __launch_bounds__(DEP_THREADS, 2) k_deposit(DParams D)
{
	constexpr int NACC = dep_nacc(WHAT), NCOMP = dep_ncomp(WHAT);
	extern __shared__ double smem[];
	double * tile = smem;                                   // [NCOMP][729]
	double * tphi = smem + NCOMP * DT_SITES;                // [729]
	__shared__ uint32_t heavy_first[GEVB_BRICK_CELLS], heavy_last[GEVB_BRICK_CELLS];
	__shared__ uint16_t heavy_cell[GEVB_BRICK_CELLS];
	__shared__ int nheavy;

	const BrickGeom & G = D.G;
	const uint32_t brick = blockIdx.x;
	const uint32_t key0 = brick * GEVB_BRICK_CELLS;
	if (D.cell_start[key0] == D.cell_start[key0 + GEVB_BRICK_CELLS]) return;   // empty brick
	int x0, y0, zl0;
	brick_origin(G, brick, x0, y0, zl0);

	for (int idx = threadIdx.x; idx < NCOMP * DT_SITES; idx += DEP_THREADS) tile[idx] = 0.;
	if (threadIdx.x == 0) nheavy = 0;
	for (int s = threadIdx.x; s < DT_SITES; s += DEP_THREADS)
	{
		double v = 0.;
		if (HAS_PHI)
		{
			const int tz = s / (DT_EDGE * DT_EDGE), ty = (s / DT_EDGE) % DT_EDGE, tx = s % DT_EDGE;
			const int plane = zl0 + tz + 1;
			if (plane <= G.nzl + 1) v = __ldg(D.phi + ((size_t) plane * G.N + (y0 + ty) % G.N) * G.N + (x0 + tx) % G.N);
		}
		tphi[s] = v;
	}
	__syncthreads();

	// ---- light pass: one thread per cell, DEP_LIGHT particles at most -------------------------
	for (int c = threadIdx.x; c < GEVB_BRICK_CELLS; c += DEP_THREADS)
	{
		const int sx = c & 7, sy = (c >> 3) & 7, sz = c >> 6;
		const int site = (sz * DT_EDGE + sy) * DT_EDGE + sx;
		const uint32_t first = D.cell_start[key0 + c], last = D.cell_start[key0 + c + 1];
		const uint32_t n = last - first;
		double acc[NACC];
		#pragma unroll
		for (int a = 0; a < NACC; a++) acc[a] = 0.;
		double cphi[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) cphi[k] = tphi[site + corner_offset(k)];       // :967-974
		if (n > 0)
		{
			const double refx = (x0 + sx) * D.dx, refy = (y0 + sy) * D.dx, refz = (G.z0 + zl0 + sz) * D.dx;   // referPos, :963
			const uint32_t nl = n < DEP_LIGHT ? n : DEP_LIGHT;
			for (uint32_t j = 0; j < nl; j++) accumulate<WHAT, HAS_PHI>(acc, D, first + j, refx, refy, refz, cphi);
			finalize<WHAT>(acc, D, cphi);
			if (n > DEP_LIGHT)
			{
				const int h = atomicAdd(&nheavy, 1);
				heavy_cell[h] = (uint16_t) c; heavy_first[h] = first + DEP_LIGHT; heavy_last[h] = last;
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
		const int sx = c & 7, sy = (c >> 3) & 7, sz = c >> 6;
		const int site = (sz * DT_EDGE + sy) * DT_EDGE + sx;
		double acc[NACC];
		#pragma unroll
		for (int a = 0; a < NACC; a++) acc[a] = 0.;
		double cphi[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) cphi[k] = tphi[site + corner_offset(k)];
		const double refx = (x0 + sx) * D.dx, refy = (y0 + sy) * D.dx, refz = (G.z0 + zl0 + sz) * D.dx;
		for (uint32_t i = heavy_first[h] + (threadIdx.x & 31); i < heavy_last[h]; i += 32) accumulate<WHAT, HAS_PHI>(acc, D, i, refx, refy, refz, cphi);
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
	__syncthreads();

	// ---- flush the tile: FP64 reductions into HBM, consecutive lanes on consecutive sites of a row
	for (int idx = threadIdx.x; idx < NCOMP * DT_SITES; idx += DEP_THREADS)
	{
		const double v = tile[idx];
		if (v == 0.) continue;
		const int comp = idx / DT_SITES, s = idx - comp * DT_SITES;
		const int tz = s / (DT_EDGE * DT_EDGE), ty = (s / DT_EDGE) % DT_EDGE, tx = s % DT_EDGE;
		const size_t off = ((size_t) (zl0 + tz + 1) * G.N + (y0 + ty) % G.N) * G.N + (x0 + tx) % G.N;
		atomicAdd(D.out[comp] + off, v);
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
	const size_t smem = (size_t) (dep_ncomp(WHAT) + 1) * DT_SITES * sizeof(double);
	if (phi)
	{
		CUDA_TRY(cudaFuncSetAttribute(k_deposit<WHAT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_deposit<WHAT, true><<<D.G.nbricks, DEP_THREADS, smem, c->stream>>>(D);
	}
	else
	{
		CUDA_TRY(cudaFuncSetAttribute(k_deposit<WHAT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		k_deposit<WHAT, false><<<D.G.nbricks, DEP_THREADS, smem, c->stream>>>(D);
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
