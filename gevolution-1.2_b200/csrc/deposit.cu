// deposit.cu -- particle -> mesh projections (CIC / staggered CIC-NGP)
//
//   projection_T00_project       gevolution.hpp:927-1022
//   projection_T0i_project       gevolution.hpp:1046-1147
//   projection_Tij_project       gevolution.hpp:1173-1297
//   scalarProjectionCIC_project  LATfield2 (main.cpp:402), plain CIC
//
// Particles are cell-sorted, one thread per particle, coalesced SoA loads
// (48 B per particle).  The phi values at the eight cell corners come through
// the read-only path (neighbouring particles share them in L1).  Contributions
// go out as FP64 reductions (RED.ADD.F64) on the target field; x and y wrap by
// index arithmetic, z+1 of the last local plane lands in the upper ghost plane
// which gevb_projection_comm folds into the next rank.
#include "gevb_internal.cuh"

namespace {

struct DGeom { int N, nzl, z0; size_t plane; double dx; };

struct Cell
{
	int x, y, zl;          // cell coordinates (zl local)
	size_t row[2][2];      // [Z][Y] -> offset of the row start, plane index zl+1+Z
	int xs[2];             // [X] -> wrapped x
};

__device__ __forceinline__ Cell cell_from_key(uint32_t key, const DGeom & G)
{
	Cell c;
	c.x = (int) (key % (uint32_t) G.N); uint32_t r = key / (uint32_t) G.N;
	c.y = (int) (r % (uint32_t) G.N); c.zl = (int) (r / (uint32_t) G.N);
	const int yp = c.y == G.N - 1 ? 0 : c.y + 1;
	c.xs[0] = c.x; c.xs[1] = c.x == G.N - 1 ? 0 : c.x + 1;
	#pragma unroll
	for (int Z = 0; Z < 2; Z++)
	{
		c.row[Z][0] = ((size_t) (c.zl + 1 + Z) * G.N + c.y) * G.N;
		c.row[Z][1] = ((size_t) (c.zl + 1 + Z) * G.N + yp) * G.N;
	}
	return c;
}

// corner index 4X + 2Y + Z (gevolution.hpp:953)
__device__ __forceinline__ size_t corner(const Cell & c, int X, int Y, int Z) { return c.row[Z][Y] + c.xs[X]; }

__device__ __forceinline__ void red_add(double * p, double v) { atomicAdd(p, v); }

template <bool DO_T00, bool DO_TIJ, bool HAS_PHI>
__global__ void __launch_bounds__(256) k_deposit_scalar_tensor(DGeom G, int64_t n, const uint32_t * __restrict__ key,
	const double * __restrict__ px, const double * __restrict__ py, const double * __restrict__ pz,
	const double * __restrict__ qx, const double * __restrict__ qy, const double * __restrict__ qz,
	const double * __restrict__ phi, double * T00, double * Tij, size_t cs, double mass, double a)
{
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
	{
		const Cell c = cell_from_key(key[i], G);
		double up[3], dn[3];
		up[0] = (px[i] - c.x * G.dx) / G.dx;                                  // gevolution.hpp:981 / :1231
		up[1] = (py[i] - c.y * G.dx) / G.dx;
		up[2] = (pz[i] - (c.zl + G.z0) * G.dx) / G.dx;                        // global cell coordinate, as xPart.coord(2)
		dn[0] = 1.0 - up[0]; dn[1] = 1.0 - up[1]; dn[2] = 1.0 - up[2];         // :982
		double cphi[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) cphi[k] = HAS_PHI ? __ldg(phi + corner(c, (k >> 2) & 1, (k >> 1) & 1, k & 1)) : 0.;
		const double q0 = qx[i], q1 = qy[i], q2 = qz[i];
		const double q2sum = q0 * q0 + q1 * q1 + q2 * q2;
		double w[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) w[k] = ((k & 4) ? up[0] : dn[0]) * ((k & 2) ? up[1] : dn[1]) * ((k & 1) ? up[2] : dn[2]);
		if (DO_T00)
		{
			double e = a, f = 0.;
			if (HAS_PHI) { e = sqrt(q2sum + a * a); f = 3. * e + q2sum / e; }    // :989-991
			#pragma unroll
			for (int k = 0; k < 8; k++)
				red_add(T00 + corner(c, (k >> 2) & 1, (k >> 1) & 1, k & 1), w[k] * (e + f * cphi[k]) * mass);   // :995-1019
		}
		if (DO_TIJ)
		{
			const double e = sqrt(q2sum + a * a);                              // :1237
			const double f = 4. + a * a / (q2sum + a * a);                     // :1238
			const double qq[3] = {q0, q1, q2};
			const int diag[3] = {0, 3, 5};
			#pragma unroll
			for (int d = 0; d < 3; d++)
			{
				const double wd = mass * qq[d] * qq[d] / e;                    // :1243
				#pragma unroll
				for (int k = 0; k < 8; k++)
					red_add(Tij + diag[d] * cs + corner(c, (k >> 2) & 1, (k >> 1) & 1, k & 1), wd * w[k] * (1. + f * cphi[k]));   // :1245-1259
			}
			double wo = mass * q0 * q1 / e;                                    // :1262-1264 -> (0,1) at x and x+e2
			red_add(Tij + 1 * cs + corner(c, 0, 0, 0), wo * dn[2] * (1. + f * 0.25 * (cphi[0] + cphi[2] + cphi[4] + cphi[6])));
			red_add(Tij + 1 * cs + corner(c, 0, 0, 1), wo * up[2] * (1. + f * 0.25 * (cphi[1] + cphi[3] + cphi[5] + cphi[7])));
			wo = mass * q0 * q2 / e;                                           // :1266-1268 -> (0,2) at x and x+e1
			red_add(Tij + 2 * cs + corner(c, 0, 0, 0), wo * dn[1] * (1. + f * 0.25 * (cphi[0] + cphi[1] + cphi[4] + cphi[5])));
			red_add(Tij + 2 * cs + corner(c, 0, 1, 0), wo * up[1] * (1. + f * 0.25 * (cphi[2] + cphi[3] + cphi[6] + cphi[7])));
			wo = mass * q1 * q2 / e;                                           // :1270-1272 -> (1,2) at x and x+e0
			red_add(Tij + 4 * cs + corner(c, 0, 0, 0), wo * dn[0] * (1. + f * 0.25 * (cphi[0] + cphi[1] + cphi[2] + cphi[3])));
			red_add(Tij + 4 * cs + corner(c, 1, 0, 0), wo * up[0] * (1. + f * 0.25 * (cphi[4] + cphi[5] + cphi[6] + cphi[7])));
		}
	}
}

template <bool HAS_PHI>
__global__ void __launch_bounds__(256) k_deposit_T0i(DGeom G, int64_t n, const uint32_t * __restrict__ key,
	const double * __restrict__ px, const double * __restrict__ py, const double * __restrict__ pz,
	const double * __restrict__ qx, const double * __restrict__ qy, const double * __restrict__ qz,
	const double * __restrict__ phi, double * T0i, size_t cs, double mass)
{
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
	{
		const Cell c = cell_from_key(key[i], G);
		double up[3], dn[3];
		up[0] = (px[i] - c.x * G.dx) / G.dx; up[1] = (py[i] - c.y * G.dx) / G.dx; up[2] = (pz[i] - (c.zl + G.z0) * G.dx) / G.dx;
		dn[0] = 1.0 - up[0]; dn[1] = 1.0 - up[1]; dn[2] = 1.0 - up[2];
		double cp[8];
		#pragma unroll
		for (int k = 0; k < 8; k++) cp[k] = HAS_PHI ? __ldg(phi + corner(c, (k >> 2) & 1, (k >> 1) & 1, k & 1)) : 0.;
		double w = mass * qx[i];                                               // :1107
		red_add(T0i + corner(c, 0, 0, 0), w * dn[1] * dn[2] * (1. + cp[0] + cp[4]));            // :1109,:1129
		red_add(T0i + corner(c, 0, 1, 0), w * up[1] * dn[2] * (1. + cp[2] + cp[6]));            // :1110,:1136
		red_add(T0i + corner(c, 0, 0, 1), w * dn[1] * up[2] * (1. + cp[1] + cp[5]));            // :1111,:1139
		red_add(T0i + corner(c, 0, 1, 1), w * up[1] * up[2] * (1. + cp[3] + cp[7]));            // :1112,:1142
		w = mass * qy[i];                                                      // :1114
		red_add(T0i + cs + corner(c, 0, 0, 0), w * dn[0] * dn[2] * (1. + cp[0] + cp[2]));       // :1116,:1130
		red_add(T0i + cs + corner(c, 1, 0, 0), w * up[0] * dn[2] * (1. + cp[4] + cp[6]));       // :1117,:1133
		red_add(T0i + cs + corner(c, 0, 0, 1), w * dn[0] * up[2] * (1. + cp[1] + cp[3]));       // :1118,:1140
		red_add(T0i + cs + corner(c, 1, 0, 1), w * up[0] * up[2] * (1. + cp[5] + cp[7]));       // :1119,:1143
		w = mass * qz[i];                                                      // :1121
		red_add(T0i + 2 * cs + corner(c, 0, 0, 0), w * dn[0] * dn[1] * (1. + cp[0] + cp[1]));   // :1123,:1131
		red_add(T0i + 2 * cs + corner(c, 1, 0, 0), w * up[0] * dn[1] * (1. + cp[4] + cp[5]));   // :1124,:1134
		red_add(T0i + 2 * cs + corner(c, 0, 1, 0), w * dn[0] * up[1] * (1. + cp[2] + cp[3]));   // :1125,:1137
		red_add(T0i + 2 * cs + corner(c, 1, 1, 0), w * up[0] * up[1] * (1. + cp[6] + cp[7]));   // :1126,:1144
	}
}

int check_real(const gevb_field * f, int ncomp, const char * who, const char * name)
{
	GEVB_CHECK_ARG(f != NULL, "%s: %s is NULL", who, name);
	GEVB_CHECK_ARG(f->kind == GEVB_REAL, "%s: %s must be a real-space field", who, name);
	GEVB_CHECK_ARG(f->ncomp == ncomp, "%s: %s needs %d components (has %d)", who, name, ncomp, f->ncomp);
	return 0;
}

template <bool DO_T00, bool DO_TIJ>
int launch_st(gevb_pcls * p, gevb_field * T00, gevb_field * Tij, double a, gevb_field * phi, double mass)
{
	gevb_ctx * c = p->ctx;
	if (p->n == 0) return 0;
	const int b = p->cur;
	DGeom G = {c->N, c->nzl, c->z0, c->plane(), 1.0 / (double) c->N};
	const int grid = gevb_grid(c, (size_t) p->n, 256);
	double * t00 = T00 ? T00->data : NULL; double * tij = Tij ? Tij->data : NULL;
	size_t cs = Tij ? Tij->comp_stride : 0;
	if (phi)
		k_deposit_scalar_tensor<DO_T00, DO_TIJ, true><<<grid, 256, 0, c->stream>>>(G, p->n, p->key[b], p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], phi->data, t00, tij, cs, mass, a);
	else
		k_deposit_scalar_tensor<DO_T00, DO_TIJ, false><<<grid, 256, 0, c->stream>>>(G, p->n, p->key[b], p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], NULL, t00, tij, cs, mass, a);
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
	return launch_st<true, false>(p, T00, NULL, a, phi, mass);
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
	return launch_st<false, true>(p, NULL, Tij, a, phi, mass);
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
	return launch_st<true, true>(p, T00, Tij, a, phi, mass);
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
	return launch_st<true, false>(p, rho, NULL, 1.0, NULL, p->mass / (dx * dx * dx));
}

extern "C" int gevb_projection_T0i_project(gevb_pcls * p, gevb_field * T0i, gevb_field * phi, double coeff)
{
	GEVB_CHECK_ARG(p != NULL, "projection_T0i_project: NULL particle handle");
	GEVB_TRY(check_real(T0i, 3, "projection_T0i_project", "T0i"));
	if (phi) GEVB_TRY(check_real(phi, 1, "projection_T0i_project", "phi"));
	gevb_ctx * c = p->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_T0I);
	if (p->n == 0) return 0;
	const double dx = 1.0 / (double) c->N;
	double mass = coeff / (dx * dx * dx); mass *= p->mass;                     // gevolution.hpp:1064-1065
	const int b = p->cur;
	DGeom G = {c->N, c->nzl, c->z0, c->plane(), dx};
	const int grid = gevb_grid(c, (size_t) p->n, 256);
	if (phi)
		k_deposit_T0i<true><<<grid, 256, 0, c->stream>>>(G, p->n, p->key[b], p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], phi->data, T0i->data, T0i->comp_stride, mass);
	else
		k_deposit_T0i<false><<<grid, 256, 0, c->stream>>>(G, p->n, p->key[b], p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], NULL, T0i->data, T0i->comp_stride, mass);
	KERNEL_CHECK(c);
	return 0;
}
