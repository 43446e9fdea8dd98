// geodesic.cu -- particle kick and drift
//
//   Particles::updateVel + update_q / update_q_Newton        main.cpp:775; gevolution.hpp:570-678, 709-776
//   Particles::moveParticles + update_pos / update_pos_Newton main.cpp:798; gevolution.hpp:810-871, 900-903
//
// One thread per particle over the cell-sorted SoA (coalesced 48 B in, 24 or
// 48 B out).  Field gathers (phi x8, chi x8, B at 36 sites) go through the
// read-only path: particles are cell-sorted, so a warp touches a handful of
// rows of each field plane and neighbouring warps re-use them from L1/L2.
// The fused kernel does kick and drift in one pass: between main.cpp:775 and
// :798 only `a` changes (rungekutta4bg, :792), positions and fields do not.
// After a drift the new cell key is written; particles that leave the z-slab
// are compacted into send buffers for the two ring neighbours (NCCL P2P).
#include <math.h>
#include "gevb_internal.cuh"

namespace {

struct GParams
{
	int N, nzl, z0, nranks;
	size_t plane, csB;
	double dx;
	const double * phi, * chi, * B;
	// kick
	int fn, nf_kick; double dtau_kick, a_kick, bscale_kick;
	// drift
	int nf_drift; double dtau_drift, a_drift, bscale_drift;
	// particles
	int64_t n;
	double * x, * y, * z, * qx, * qy, * qz; int64_t * id; uint32_t * key;
	unsigned long long * maxv2;          // bit pattern of the running max of v^2 (>= 0)
	// migration
	unsigned long long * nsend;          // [2]: down, up
	double * sendbuf[2]; int64_t sendcap; uint32_t invalid_key;
};

struct Stencil
{
	int xi[3], yi[3];      // wrapped x-1,x,x+1 ; y-1,y,y+1
	int p;                 // plane index of the particle's cell (local z + 1)
	int N; size_t plane;
	__device__ __forceinline__ size_t at(int dx, int dy, int dz) const { return ((size_t) (p + dz) * N + yi[dy + 1]) * N + xi[dx + 1]; }
};

__device__ __forceinline__ int cell_of(double p, double dx, int N)
{
	int c = (int) floor(p / dx);
	c = c >= N ? N - 1 : c;
	return c < 0 ? 0 : c;
}

// one-sided CIC gradient, gevolution.hpp:585-596 (GRADIENT_ORDER == 1)
__device__ __forceinline__ void grad_cic(const double * __restrict__ f, const Stencil & s, const double * r, double * g)
{
	const double f000 = __ldg(f + s.at(0, 0, 0)), f100 = __ldg(f + s.at(1, 0, 0)), f010 = __ldg(f + s.at(0, 1, 0)), f110 = __ldg(f + s.at(1, 1, 0));
	const double f001 = __ldg(f + s.at(0, 0, 1)), f101 = __ldg(f + s.at(1, 0, 1)), f011 = __ldg(f + s.at(0, 1, 1)), f111 = __ldg(f + s.at(1, 1, 1));
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
__device__ __forceinline__ double tri_cic(const double * __restrict__ f, const Stencil & s, const double * r)
{
	double v = __ldg(f + s.at(0, 0, 0)) * (1. - r[0]) * (1. - r[1]) * (1. - r[2]);
	v += __ldg(f + s.at(1, 0, 0)) * r[0] * (1. - r[1]) * (1. - r[2]);
	v += __ldg(f + s.at(0, 1, 0)) * (1. - r[0]) * r[1] * (1. - r[2]);
	v += __ldg(f + s.at(1, 1, 0)) * r[0] * r[1] * (1. - r[2]);
	v += __ldg(f + s.at(0, 0, 1)) * (1. - r[0]) * (1. - r[1]) * r[2];
	v += __ldg(f + s.at(1, 0, 1)) * r[0] * (1. - r[1]) * r[2];
	v += __ldg(f + s.at(0, 1, 1)) * (1. - r[0]) * r[1] * r[2];
	v += __ldg(f + s.at(1, 1, 1)) * r[0] * r[1] * r[2];
	return v;
}

// update_q (gevolution.hpp:570-678) / update_q_Newton (:709-776); returns v^2/a^2
__device__ __forceinline__ double kick(const GParams & P, const Stencil & s, const double * r, double * q)
{
	double g[3], v2;
	if (P.fn == GEVB_UPDATE_Q)
	{
		v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                             // :581
		double e2 = v2 + P.a_kick * P.a_kick;                                      // :582
		grad_cic(P.phi, s, r, g);                                                  // :585-596
		g[0] *= (v2 + e2) / e2; g[1] *= (v2 + e2) / e2; g[2] *= (v2 + e2) / e2;    // :613-615
		if (P.nf_kick >= 2 && P.chi != NULL)
		{
			double gc[3]; grad_cic(P.chi, s, r, gc);                               // :617-631
			g[0] -= gc[0]; g[1] -= gc[1]; g[2] -= gc[2];
		}
		e2 = sqrt(e2);                                                             // :633
		if (P.nf_kick >= 3 && P.B != NULL)
		{
			const double * B0 = P.B, * B1 = P.B + P.csB, * B2 = P.B + 2 * P.csB;
#define BV(Bc, a_, b_, c_) __ldg((Bc) + s.at(a_, b_, c_))
			double pg0, pg1, pg2;
			// :637-642
			pg0 = ((1. - r[2]) * (BV(B1, 1, 0, 0) - BV(B1, 0, 0, 0)) + r[2] * (BV(B1, 1, 0, 1) - BV(B1, 0, 0, 1))) * q[1];
			pg0 += ((1. - r[1]) * (BV(B2, 1, 0, 0) - BV(B2, 0, 0, 0)) + r[1] * (BV(B2, 1, 1, 0) - BV(B2, 0, 1, 0))) * q[2];
			pg0 += (1. - r[1]) * (1. - r[2]) * ((r[0] - 1.) * BV(B0, -1, 0, 0) + (1. - 2. * r[0]) * BV(B0, 0, 0, 0) + r[0] * BV(B0, 1, 0, 0)) * q[0];
			pg0 += r[1] * (1. - r[2]) * ((r[0] - 1.) * BV(B0, -1, 1, 0) + (1. - 2. * r[0]) * BV(B0, 0, 1, 0) + r[0] * BV(B0, 1, 1, 0)) * q[0];
			pg0 += (1. - r[1]) * r[2] * ((r[0] - 1.) * BV(B0, -1, 0, 1) + (1. - 2. * r[0]) * BV(B0, 0, 0, 1) + r[0] * BV(B0, 1, 0, 1)) * q[0];
			pg0 += r[1] * r[2] * ((r[0] - 1.) * BV(B0, -1, 1, 1) + (1. - 2. * r[0]) * BV(B0, 0, 1, 1) + r[0] * BV(B0, 1, 1, 1)) * q[0];
			// :644-649
			pg1 = ((1. - r[0]) * (BV(B2, 0, 1, 0) - BV(B2, 0, 0, 0)) + r[0] * (BV(B2, 1, 1, 0) - BV(B2, 1, 0, 0))) * q[2];
			pg1 += ((1. - r[2]) * (BV(B0, 0, 1, 0) - BV(B0, 0, 0, 0)) + r[2] * (BV(B0, 0, 1, 1) - BV(B0, 0, 0, 1))) * q[0];
			pg1 += (1. - r[0]) * (1. - r[2]) * ((r[1] - 1.) * BV(B1, 0, -1, 0) + (1. - 2. * r[1]) * BV(B1, 0, 0, 0) + r[1] * BV(B1, 0, 1, 0)) * q[1];
			pg1 += r[0] * (1. - r[2]) * ((r[1] - 1.) * BV(B1, 1, -1, 0) + (1. - 2. * r[1]) * BV(B1, 1, 0, 0) + r[1] * BV(B1, 1, 1, 0)) * q[1];
			pg1 += (1. - r[0]) * r[2] * ((r[1] - 1.) * BV(B1, 0, -1, 1) + (1. - 2. * r[1]) * BV(B1, 0, 0, 1) + r[1] * BV(B1, 0, 1, 1)) * q[1];
			pg1 += r[0] * r[2] * ((r[1] - 1.) * BV(B1, 1, -1, 1) + (1. - 2. * r[1]) * BV(B1, 1, 0, 1) + r[1] * BV(B1, 1, 1, 1)) * q[1];
			// :651-656
			pg2 = ((1. - r[1]) * (BV(B0, 0, 0, 1) - BV(B0, 0, 0, 0)) + r[1] * (BV(B0, 0, 1, 1) - BV(B0, 0, 1, 0))) * q[0];
			pg2 += ((1. - r[0]) * (BV(B1, 0, 0, 1) - BV(B1, 0, 0, 0)) + r[0] * (BV(B1, 1, 0, 1) - BV(B1, 1, 0, 0))) * q[1];
			pg2 += (1. - r[0]) * (1. - r[1]) * ((r[2] - 1.) * BV(B2, 0, 0, -1) + (1. - 2. * r[2]) * BV(B2, 0, 0, 0) + r[2] * BV(B2, 0, 0, 1)) * q[2];
			pg2 += r[0] * (1. - r[1]) * ((r[2] - 1.) * BV(B2, 1, 0, -1) + (1. - 2. * r[2]) * BV(B2, 1, 0, 0) + r[2] * BV(B2, 1, 0, 1)) * q[2];
			pg2 += (1. - r[0]) * r[1] * ((r[2] - 1.) * BV(B2, 0, 1, -1) + (1. - 2. * r[2]) * BV(B2, 0, 1, 0) + r[2] * BV(B2, 0, 1, 1)) * q[2];
			pg2 += r[0] * r[1] * ((r[2] - 1.) * BV(B2, 1, 1, -1) + (1. - 2. * r[2]) * BV(B2, 1, 1, 0) + r[2] * BV(B2, 1, 1, 1)) * q[2];
#undef BV
			g[0] += pg0 / P.bscale_kick / e2;                                      // :658-660
			g[1] += pg1 / P.bscale_kick / e2;
			g[2] += pg2 / P.bscale_kick / e2;
		}
		v2 = 0.;
		#pragma unroll
		for (int i = 0; i < 3; i++) { q[i] -= P.dtau_kick * e2 * g[i] / P.dx; v2 += q[i] * q[i]; }   // :664-668
	}
	else
	{
		grad_cic(P.phi, s, r, g);                                                  // :719-730
		if (P.nf_kick >= 2 && P.chi != NULL)
		{
			double gc[3]; grad_cic(P.chi, s, r, gc);                               // :747-761
			g[0] -= gc[0]; g[1] -= gc[1]; g[2] -= gc[2];
		}
		v2 = 0.;
		#pragma unroll
		for (int i = 0; i < 3; i++) { q[i] -= P.dtau_kick * P.a_kick * g[i] / P.dx; v2 += q[i] * q[i]; }   // :764-768
	}
	return v2 / P.a_kick / P.a_kick;                                               // :670 / :770
}

// update_pos (gevolution.hpp:810-871) / update_pos_Newton (:900-903)
__device__ __forceinline__ void drift(const GParams & P, const Stencil & s, const double * r, const double * q, double * pos)
{
	if (P.fn != GEVB_UPDATE_Q)
	{
		#pragma unroll
		for (int l = 0; l < 3; l++) pos[l] += P.dtau_drift * q[l] / P.a_drift;     // :902
		return;
	}
	double v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                           // :813
	const double e2 = v2 + P.a_drift * P.a_drift;                                  // :814
	double ph = 0., ch = 0.;
	if (P.nf_drift >= 1) ph = tri_cic(P.phi, s, r);                                // :820-827
	if (P.nf_drift >= 2) ch = tri_cic(P.chi, s, r);                                // :832-839
	v2 = (1. + (3. - v2 / e2) * ph - ch) / sqrt(e2);                               // :842
	double v[3] = {q[0] * v2, q[1] * v2, q[2] * v2};                               // :844-846
	if (P.nf_drift >= 3)
	{
		const double * B0 = P.B, * B1 = P.B + P.csB, * B2 = P.B + 2 * P.csB;
#define BV(Bc, a_, b_, c_) __ldg((Bc) + s.at(a_, b_, c_))
		double b[3];
		b[0] = BV(B0, 0, 0, 0) * (1. - r[1]) * (1. - r[2]);                        // :852
		b[1] = BV(B1, 0, 0, 0) * (1. - r[0]) * (1. - r[2]);                        // :853
		b[2] = BV(B2, 0, 0, 0) * (1. - r[0]) * (1. - r[1]);                        // :854
		b[1] += BV(B1, 1, 0, 0) * r[0] * (1. - r[2]);                              // :855
		b[2] += BV(B2, 1, 0, 0) * r[0] * (1. - r[1]);                              // :856
		b[0] += BV(B0, 0, 1, 0) * r[1] * (1. - r[2]);                              // :857
		b[2] += BV(B2, 0, 1, 0) * (1. - r[0]) * r[1];                              // :858
		b[0] += BV(B0, 0, 0, 1) * (1. - r[1]) * r[2];                              // :859
		b[1] += BV(B1, 0, 0, 1) * (1. - r[0]) * r[2];                              // :860
		b[1] += BV(B1, 1, 0, 1) * r[0] * r[2];                                     // :861
		b[0] += BV(B0, 0, 1, 1) * r[1] * r[2];                                     // :862
		b[2] += BV(B2, 1, 1, 0) * r[0] * r[1];                                     // :863
#undef BV
		#pragma unroll
		for (int l = 0; l < 3; l++) pos[l] += P.dtau_drift * (v[l] + b[l] / P.bscale_drift);   // :865
	}
	else
	{
		#pragma unroll
		for (int l = 0; l < 3; l++) pos[l] += P.dtau_drift * v[l];                 // :869
	}
}

// periodic wrap into [0,1): p - floor(p), a result that rounds to 1 maps to 0 (edge semantics)
__device__ __forceinline__ double wrap_pos(double p)
{
	double w = p - floor(p);
	return w >= 1.0 ? 0. : w;
}

// MODE 0: kick only, 1: drift only, 2: fused kick + drift
template <int MODE>
__global__ void __launch_bounds__(256) k_geodesic(GParams P)
{
	double vmax = 0.;
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < P.n; i += (int64_t) gridDim.x * blockDim.x)
	{
		double pos[3] = {P.x[i], P.y[i], P.z[i]};
		double q[3] = {P.qx[i], P.qy[i], P.qz[i]};
		// the particle's cell and ref_dist = frac(pos/dx) (LATfield2 updateVel / moveParticles drivers)
		double r[3], ip;
		Stencil s;
		s.N = P.N; s.plane = P.plane;
		{
			const int cx = cell_of(pos[0], P.dx, P.N), cy = cell_of(pos[1], P.dx, P.N), cz = cell_of(pos[2], P.dx, P.N);
			s.xi[1] = cx; s.xi[0] = cx == 0 ? P.N - 1 : cx - 1; s.xi[2] = cx == P.N - 1 ? 0 : cx + 1;
			s.yi[1] = cy; s.yi[0] = cy == 0 ? P.N - 1 : cy - 1; s.yi[2] = cy == P.N - 1 ? 0 : cy + 1;
			s.p = cz - P.z0 + 1;
			r[0] = modf(pos[0] / P.dx, &ip); r[1] = modf(pos[1] / P.dx, &ip); r[2] = modf(pos[2] / P.dx, &ip);
		}
		if (MODE == 0 || MODE == 2)
		{
			const double v2 = kick(P, s, r, q);
			vmax = fmax(vmax, v2);
			P.qx[i] = q[0]; P.qy[i] = q[1]; P.qz[i] = q[2];
		}
		if (MODE == 1 || MODE == 2)
		{
			drift(P, s, r, q, pos);
			pos[0] = wrap_pos(pos[0]); pos[1] = wrap_pos(pos[1]); pos[2] = wrap_pos(pos[2]);
			const int cx = cell_of(pos[0], P.dx, P.N), cy = cell_of(pos[1], P.dx, P.N), cz = cell_of(pos[2], P.dx, P.N);
			int zl = cz - P.z0;
			bool leaving = false;
			if (P.nranks > 1 && (zl < 0 || zl >= P.nzl))
			{
				// at most one slab per move (main.cpp:281-286): periodic distance decides the neighbour
				const int d = (zl + P.N) % P.N;
				const int dir = d < P.N / 2 ? 1 : 0;
				const unsigned long long slot = atomicAdd(P.nsend + dir, 1ull);
				if ((int64_t) slot < P.sendcap)
				{
					double * sb = P.sendbuf[dir];
					sb[slot] = pos[0]; sb[P.sendcap + slot] = pos[1]; sb[2 * P.sendcap + slot] = pos[2];
					sb[3 * P.sendcap + slot] = q[0]; sb[4 * P.sendcap + slot] = q[1]; sb[5 * P.sendcap + slot] = q[2];
					sb[6 * P.sendcap + slot] = __longlong_as_double((long long) P.id[i]);
				}
				leaving = true;
			}
			P.x[i] = pos[0]; P.y[i] = pos[1]; P.z[i] = pos[2];
			P.key[i] = leaving ? P.invalid_key : (uint32_t) ((zl * P.N + cy) * P.N + cx);
		}
	}
	if (MODE == 0 || MODE == 2)
	{
		for (int o = 16; o > 0; o >>= 1) vmax = fmax(vmax, __shfl_down_sync(0xffffffffu, vmax, o));
		if ((threadIdx.x & 31) == 0 && vmax > 0.) atomicMax(P.maxv2, (unsigned long long) __double_as_longlong(vmax));
	}
}

// received particles (7 x cap SoA staging) appended behind the live ones, keys computed
__global__ void k_append_received(int64_t nrecv, const double * __restrict__ rb, int64_t cap, int64_t at,
                                  double * x, double * y, double * z, double * qx, double * qy, double * qz, int64_t * id, uint32_t * key,
                                  int N, int z0, int nzl, double dx, unsigned long long * lost)
{
	for (int64_t i = blockIdx.x * (int64_t) blockDim.x + threadIdx.x; i < nrecv; i += (int64_t) gridDim.x * blockDim.x)
	{
		const double px = rb[i], py = rb[cap + i], pz = rb[2 * cap + i];
		x[at + i] = px; y[at + i] = py; z[at + i] = pz;
		qx[at + i] = rb[3 * cap + i]; qy[at + i] = rb[4 * cap + i]; qz[at + i] = rb[5 * cap + i];
		id[at + i] = (int64_t) __double_as_longlong(rb[6 * cap + i]);
		const int cx = cell_of(px, dx, N), cy = cell_of(py, dx, N);
		int cz = cell_of(pz, dx, N) - z0;
		if (cz < 0 || cz >= nzl) { atomicAdd(lost, 1ull); cz = cz < 0 ? 0 : nzl - 1; }   // moved farther than one slab: reported by the host
		key[at + i] = (uint32_t) ((cz * N + cy) * N + cx);
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
	P.N = c->N; P.nzl = c->nzl; P.z0 = c->z0; P.nranks = c->nranks;
	P.plane = c->plane(); P.dx = 1.0 / (double) c->N;
	P.phi = nfields >= 1 ? fields[0]->data : NULL;
	P.chi = nfields >= 2 ? fields[1]->data : NULL;
	P.B = nfields >= 3 ? fields[2]->data : NULL;
	P.csB = nfields >= 3 ? fields[2]->comp_stride : 0;
	P.n = p->n;
	P.x = p->x[b]; P.y = p->y[b]; P.z = p->z[b]; P.qx = p->qx[b]; P.qy = p->qy[b]; P.qz = p->qz[b]; P.id = p->id[b]; P.key = p->key[b];
	P.maxv2 = (unsigned long long *) (c->d_red + 4008);
	P.nsend = (unsigned long long *) (c->d_red + 4010);
	P.invalid_key = 0xffffffffu;
}

// after a drift: exchange slab-crossing particles with the ring neighbours, then restore the sort
int finish_move(gevb_pcls * p, GParams & P)
{
	gevb_ctx * c = p->ctx;
	if (c->nranks == 1) { Timed timed_(c, CLS_SORT); return gevb_pcls_sort(p, true, false); }
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
		k_append_received<<<gevb_grid(c, h[2], 256), 256, 0, c->stream>>>((int64_t) h[2], rb_up, P.sendcap, p->n, p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], p->key[b], c->N, c->z0, c->nzl, dx, cnt + 4);
		KERNEL_CHECK(c);
	}
	if (h[3])
	{
		k_append_received<<<gevb_grid(c, h[3], 256), 256, 0, c->stream>>>((int64_t) h[3], rb_dn, P.sendcap, p->n + (int64_t) h[2], p->x[b], p->y[b], p->z[b], p->qx[b], p->qy[b], p->qz[b], p->id[b], p->key[b], c->N, c->z0, c->nzl, dx, cnt + 4);
		KERNEL_CHECK(c);
	}
	p->n += nrecv;
	// keys of the departed are 0xffffffff: they sort to the very end and are dropped
	GEVB_TRY(gevb_pcls_sort(p, true, true));
	p->n -= (int64_t) (h[0] + h[1]);
	unsigned long long lost = 0;
	CUDA_TRY(cudaMemcpyAsync(&lost, cnt + 4, sizeof(lost), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	GEVB_CHECK_ARG(lost == 0, "moveParticles: %llu particles moved farther than the adjacent slab (move limit, main.cpp:281-286)", lost);
	return 0;
}

int setup_migration(gevb_pcls * p, GParams & P)
{
	gevb_ctx * c = p->ctx;
	if (c->nranks == 1) return 0;
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
	GEVB_CHECK_ARG(p != NULL && params != NULL, "updateVel: NULL argument");
	GEVB_CHECK_ARG(fn == GEVB_UPDATE_Q || fn == GEVB_UPDATE_Q_NEWTON, "updateVel: unknown callback %d (only update_q and update_q_Newton can run on the device)", fn);
	gevb_ctx * c = p->ctx;
	GEVB_TRY(check_fields(c, fields, nfields, "updateVel"));
	GEVB_CHECK_ARG(nfields >= 1, "updateVel: needs at least phi");
	CUDA_TRY(cudaSetDevice(c->device));
	GParams P;
	base_params(P, p, fields, nfields);
	P.fn = fn; P.nf_kick = nfields; P.dtau_kick = dtau; P.a_kick = params[0]; P.bscale_kick = params[1];
	CUDA_TRY(cudaMemsetAsync(P.maxv2, 0, sizeof(unsigned long long), c->stream));
	if (p->n > 0)
	{
		Timed timed_(c, CLS_KICK);
		k_geodesic<0><<<gevb_grid(c, (size_t) p->n, 256), 256, 0, c->stream>>>(P);
		KERNEL_CHECK(c);
	}
	if (maxvel)
	{
		CUDA_TRY(cudaMemcpyAsync(c->h_red, P.maxv2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*maxvel = sqrt(c->h_red[0]);           // LATfield2 updateVel returns sqrt(max v^2), cf. main.cpp:816-822
	}
	return 0;
}

extern "C" int gevb_moveParticles(gevb_pcls * p, int fn, double dtau, gevb_field * const * fields, int nfields, const double * params)
{
	GEVB_CHECK_ARG(p != NULL && params != NULL, "moveParticles: NULL argument");
	GEVB_CHECK_ARG(fn == GEVB_UPDATE_Q || fn == GEVB_UPDATE_Q_NEWTON, "moveParticles: unknown callback %d (only update_pos and update_pos_Newton can run on the device)", fn);
	gevb_ctx * c = p->ctx;
	if (fn == GEVB_UPDATE_Q_NEWTON) nfields = 0;
	GEVB_TRY(check_fields(c, fields, nfields, "moveParticles"));
	CUDA_TRY(cudaSetDevice(c->device));
	GParams P;
	base_params(P, p, fields, nfields);
	P.fn = fn; P.nf_drift = nfields; P.dtau_drift = dtau; P.a_drift = params[0]; P.bscale_drift = params[1];
	GEVB_TRY(setup_migration(p, P));
	if (p->n > 0)
	{
		Timed timed_(c, CLS_DRIFT);
		k_geodesic<1><<<gevb_grid(c, (size_t) p->n, 256), 256, 0, c->stream>>>(P);
		KERNEL_CHECK(c);
	}
	return finish_move(p, P);
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
	P.nf_kick = nfields_kick; P.dtau_kick = dtau_kick; P.a_kick = params_kick[0]; P.bscale_kick = params_kick[1];
	P.nf_drift = nfields_drift; P.dtau_drift = dtau_drift; P.a_drift = params_drift[0]; P.bscale_drift = params_drift[1];
	GEVB_TRY(setup_migration(p, P));
	CUDA_TRY(cudaMemsetAsync(P.maxv2, 0, sizeof(unsigned long long), c->stream));
	if (p->n > 0)
	{
		Timed timed_(c, CLS_KICK_DRIFT);
		k_geodesic<2><<<gevb_grid(c, (size_t) p->n, 256), 256, 0, c->stream>>>(P);
		KERNEL_CHECK(c);
	}
	// the max must be read before finish_move reuses the reduction slots' neighbourhood; slots are distinct
	GEVB_TRY(finish_move(p, P));
	if (maxvel)
	{
		CUDA_TRY(cudaMemcpyAsync(c->h_red, P.maxv2, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*maxvel = sqrt(c->h_red[0]);
	}
	return 0;
}
