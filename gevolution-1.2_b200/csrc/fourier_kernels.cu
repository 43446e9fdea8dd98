// fourier_kernels.cu -- pointwise Fourier-space kernels, one HBM pass each
//
//   solveModifiedPoissonFT  gevolution.hpp:501-535   32 B / k-site
//   projectFTscalar         gevolution.hpp:211-265  112 B / k-site
//   evolveFTvector          gevolution.hpp:284-330  192 B / k-site
//   projectFTvector         gevolution.hpp:350-392   96 B / k-site
//   projectFTtensor         gevolution.hpp:411-482  192 B / k-site
//
// All are HBM-bound streaming kernels: one thread per k-site, 16-byte loads
// coalesced along the fastest lattice index, k tables (N entries) read through
// L1.  Outputs may alias inputs: every site reads all its inputs before writing.
#include "gevb_internal.cuh"

namespace {

__device__ __forceinline__ double2 cmk(double r, double i) { return make_double2(r, i); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return cmk(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return cmk(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return cmk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cscale(double2 a, double s) { return cmk(a.x * s, a.y * s); }
__device__ __forceinline__ double2 cdiv(double2 a, double s) { return cmk(a.x / s, a.y / s); }
__device__ __forceinline__ double2 cconj(double2 a) { return cmk(a.x, -a.y); }

// streaming 16-byte accesses; data is touched once per kernel
__device__ __forceinline__ double2 ldc(const double2 * p) { return __ldcs(p); }
__device__ __forceinline__ void stc(double2 * p, double2 v) { __stcs(p, v); }

#define K_SITE_LOOP(L) for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < (L).sites; i += (size_t) gridDim.x * blockDim.x)

__global__ void __launch_bounds__(256) k_poisson(KLayout L, const double * __restrict__ gridk2, const double2 * src, double2 * pot, double coeff, double modif)
{
	K_SITE_LOOP(L)
	{
		int kx, ky, kz; k_decode(L, i, kx, ky, kz);
		double2 s = ldc(src + i), r;
		if ((kx | ky | kz) == 0) r = (modif == 0.) ? cmk(0., 0.) : cdiv(cscale(s, coeff), modif);          // gevolution.hpp:519-527
		else r = cdiv(cscale(s, coeff), __ldg(gridk2 + kx) + __ldg(gridk2 + ky) + __ldg(gridk2 + kz) + modif);   // :531
		stc(pot + i, r);
	}
}

__global__ void __launch_bounds__(256) k_ftscalar(KLayout L, const double * __restrict__ gridk2, const double2 * __restrict__ kshift, const double2 * S, size_t cs, double2 * chi, int add)
{
	K_SITE_LOOP(L)
	{
		int kx, ky, kz; k_decode(L, i, kx, ky, kz);
		if ((kx | ky | kz) == 0) { stc(chi + i, cmk(0., 0.)); continue; }                                   // :230-234
		const double g0 = __ldg(gridk2 + kx), g1 = __ldg(gridk2 + ky), g2 = __ldg(gridk2 + kz);
		const double2 k0 = __ldg(kshift + kx), k1 = __ldg(kshift + ky), k2 = __ldg(kshift + kz);
		double2 num = cscale(ldc(S + i), g1 + g2 - 2. * g0);                                                 // :253
		num = cadd(num, cscale(ldc(S + 3 * cs + i), g0 + g2 - 2. * g1));                                     // :254
		num = cadd(num, cscale(ldc(S + 5 * cs + i), g0 + g1 - 2. * g2));                                     // :255
		num = csub(num, cmul(cmul(cscale(k0, 6.), k1), ldc(S + 1 * cs + i)));                                // :256
		num = csub(num, cmul(cmul(cscale(k0, 6.), k2), ldc(S + 2 * cs + i)));                                // :257
		num = csub(num, cmul(cmul(cscale(k1, 6.), k2), ldc(S + 4 * cs + i)));                                // :258
		double2 r = cdiv(num, 2. * (g0 + g1 + g2) * (g0 + g1 + g2) * L.N);                                   // :259
		if (add) r = cadd(ldc(chi + i), r);                                                                  // :240
		stc(chi + i, r);
	}
}

__global__ void __launch_bounds__(256) k_evolve_vector(KLayout L, const double * __restrict__ gridk2, const double2 * __restrict__ kshift, const double2 * S, size_t cs, double2 * B, size_t cb, double a2dtau)
{
	K_SITE_LOOP(L)
	{
		int kx, ky, kz; k_decode(L, i, kx, ky, kz);
		if ((kx | ky | kz) == 0)                                                                               // :304-310
		{
			stc(B + i, cmk(0., 0.)); stc(B + cb + i, cmk(0., 0.)); stc(B + 2 * cb + i, cmk(0., 0.));
			continue;
		}
		const double g0 = __ldg(gridk2 + kx), g1 = __ldg(gridk2 + ky), g2 = __ldg(gridk2 + kz);
		const double2 k0 = __ldg(kshift + kx), k1 = __ldg(kshift + ky), k2 = __ldg(kshift + kz);
		double k4 = g0 + g1 + g2; k4 *= k4;                                                                  // :314-315
		const double2 pref = cmk(0., -2. * a2dtau / k4);
		const double2 S00 = ldc(S + i), S01 = ldc(S + cs + i), S02 = ldc(S + 2 * cs + i);
		const double2 S11 = ldc(S + 3 * cs + i), S12 = ldc(S + 4 * cs + i), S22 = ldc(S + 5 * cs + i);
		double2 t1, t2;
		// :317-319
		t1 = csub(csub(csub(cscale(S00, g1 + g2), cscale(S11, g1)), cscale(S22, g2)), cmul(cmul(cscale(k1, 2.), k2), S12));
		t2 = cscale(cadd(cmul(k1, S01), cmul(k2, S02)), g1 + g2 - g0);
		const double2 b0 = cadd(ldc(B + i), cmul(pref, cadd(cmul(cconj(k0), t1), t2)));
		stc(B + i, b0);
		// :320-322
		t1 = csub(csub(csub(cscale(S11, g0 + g2), cscale(S00, g0)), cscale(S22, g2)), cmul(cmul(cscale(k0, 2.), k2), S02));
		t2 = cscale(cadd(cmul(k0, S01), cmul(k2, S12)), g0 + g2 - g1);
		const double2 b1 = cadd(ldc(B + cb + i), cmul(pref, cadd(cmul(cconj(k1), t1), t2)));
		stc(B + cb + i, b1);
		// :323-325
		t1 = csub(csub(csub(cscale(S22, g0 + g1), cscale(S00, g0)), cscale(S11, g1)), cmul(cmul(cscale(k0, 2.), k1), S01));
		t2 = cscale(cadd(cmul(k0, S02), cmul(k1, S12)), g0 + g1 - g2);
		const double2 b2 = cadd(ldc(B + 2 * cb + i), cmul(pref, cadd(cmul(cconj(k2), t1), t2)));
		stc(B + 2 * cb + i, b2);
	}
}

// projectFTscalar (gevolution.hpp:211-265) and evolveFTvector (:284-330) on the same read of the six S components:
// main.cpp:558 and :586 both consume SijFT and write different fields (the backward transform of chi in between does
// not touch it), so one pass over SijFT serves both -- 208 instead of 304 bytes per k-site
__global__ void __launch_bounds__(256) k_ftscalar_evolve(KLayout L, const double * __restrict__ gridk2, const double2 * __restrict__ kshift, const double2 * S, size_t cs,
                                                         double2 * chi, double2 * B, size_t cb, double a2dtau)
{
	K_SITE_LOOP(L)
	{
		int kx, ky, kz; k_decode(L, i, kx, ky, kz);
		if ((kx | ky | kz) == 0)
		{
			stc(chi + i, cmk(0., 0.));                                                                           // :230-234
			stc(B + i, cmk(0., 0.)); stc(B + cb + i, cmk(0., 0.)); stc(B + 2 * cb + i, cmk(0., 0.));             // :304-310
			continue;
		}
		const double g0 = __ldg(gridk2 + kx), g1 = __ldg(gridk2 + ky), g2 = __ldg(gridk2 + kz);
		const double2 k0 = __ldg(kshift + kx), k1 = __ldg(kshift + ky), k2 = __ldg(kshift + kz);
		const double2 S00 = ldc(S + i), S01 = ldc(S + cs + i), S02 = ldc(S + 2 * cs + i);
		const double2 S11 = ldc(S + 3 * cs + i), S12 = ldc(S + 4 * cs + i), S22 = ldc(S + 5 * cs + i);
		// ---- chi, same operation order as k_ftscalar
		double2 num = cscale(S00, g1 + g2 - 2. * g0);                                                            // :253
		num = cadd(num, cscale(S11, g0 + g2 - 2. * g1));                                                         // :254
		num = cadd(num, cscale(S22, g0 + g1 - 2. * g2));                                                         // :255
		num = csub(num, cmul(cmul(cscale(k0, 6.), k1), S01));                                                    // :256
		num = csub(num, cmul(cmul(cscale(k0, 6.), k2), S02));                                                    // :257
		num = csub(num, cmul(cmul(cscale(k1, 6.), k2), S12));                                                    // :258
		stc(chi + i, cdiv(num, 2. * (g0 + g1 + g2) * (g0 + g1 + g2) * L.N));                                     // :259
		// ---- B, same operation order as k_evolve_vector
		double k4 = g0 + g1 + g2; k4 *= k4;                                                                      // :314-315
		const double2 pref = cmk(0., -2. * a2dtau / k4);
		double2 t1, t2;
		t1 = csub(csub(csub(cscale(S00, g1 + g2), cscale(S11, g1)), cscale(S22, g2)), cmul(cmul(cscale(k1, 2.), k2), S12));
		t2 = cscale(cadd(cmul(k1, S01), cmul(k2, S02)), g1 + g2 - g0);
		stc(B + i, cadd(ldc(B + i), cmul(pref, cadd(cmul(cconj(k0), t1), t2))));                                 // :317-319
		t1 = csub(csub(csub(cscale(S11, g0 + g2), cscale(S00, g0)), cscale(S22, g2)), cmul(cmul(cscale(k0, 2.), k2), S02));
		t2 = cscale(cadd(cmul(k0, S01), cmul(k2, S12)), g0 + g2 - g1);
		stc(B + cb + i, cadd(ldc(B + cb + i), cmul(pref, cadd(cmul(cconj(k1), t1), t2))));                       // :320-322
		t1 = csub(csub(csub(cscale(S22, g0 + g1), cscale(S00, g0)), cscale(S11, g1)), cmul(cmul(cscale(k0, 2.), k1), S01));
		t2 = cscale(cadd(cmul(k0, S02), cmul(k1, S12)), g0 + g1 - g2);
		stc(B + 2 * cb + i, cadd(ldc(B + 2 * cb + i), cmul(pref, cadd(cmul(cconj(k2), t1), t2))));               // :323-325
	}
}

__global__ void __launch_bounds__(256) k_ftvector(KLayout L, const double * __restrict__ gridk2, const double2 * __restrict__ kshift, const double2 * Si, double2 * B, size_t cs, double coeff, double modif)
{
	K_SITE_LOOP(L)
	{
		int kx, ky, kz; k_decode(L, i, kx, ky, kz);
		if ((kx | ky | kz) == 0) { stc(B + i, cmk(0., 0.)); stc(B + cs + i, cmk(0., 0.)); stc(B + 2 * cs + i, cmk(0., 0.)); continue; }   // :371-377
		const double kk2 = __ldg(gridk2 + kx) + __ldg(gridk2 + ky) + __ldg(gridk2 + kz);                     // :381
		const double2 k0 = __ldg(kshift + kx), k1 = __ldg(kshift + ky), k2 = __ldg(kshift + kz);
		const double2 s0 = ldc(Si + i), s1 = ldc(Si + cs + i), s2 = ldc(Si + 2 * cs + i);
		const double2 tmp = cdiv(cadd(cadd(cmul(k0, s0), cmul(k1, s1)), cmul(k2, s2)), kk2);                 // :383
		stc(B + i, cdiv(cscale(cscale(csub(s0, cmul(cconj(k0), tmp)), 4.), coeff), kk2 + modif));            // :385
		stc(B + cs + i, cdiv(cscale(cscale(csub(s1, cmul(cconj(k1), tmp)), 4.), coeff), kk2 + modif));       // :386
		stc(B + 2 * cs + i, cdiv(cscale(cscale(csub(s2, cmul(cconj(k2), tmp)), 4.), coeff), kk2 + modif));   // :387
	}
}

__global__ void __launch_bounds__(256) k_fttensor(KLayout L, const double * __restrict__ gridk2, const double2 * __restrict__ kshift, const double2 * S, double2 * h, size_t cs)
{
	K_SITE_LOOP(L)
	{
		int kc[3]; k_decode(L, i, kc[0], kc[1], kc[2]);
		if ((kc[0] | kc[1] | kc[2]) == 0) { for (int c = 0; c < 6; c++) stc(h + c * cs + i, cmk(0., 0.)); continue; }   // :432-438
		double g[3]; double2 kk[3];
		for (int d = 0; d < 3; d++) { g[d] = __ldg(gridk2 + kc[d]); kk[d] = __ldg(kshift + kc[d]); }
		double2 Sl[6];
		for (int c = 0; c < 6; c++) Sl[c] = ldc(S + c * cs + i);                                             // :442-447
		const double kk2 = g[0] + g[1] + g[2];                                                               // :449
		const double k6 = kk2 * kk2 * kk2 * L.N;                                                             // :450
		const int diag[3] = {0, 3, 5};
		const int off[3][3] = {{-1, 1, 2}, {1, -1, 4}, {2, 4, -1}};
		#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			// diagonal (a,a): :452-455, :465-468, :474-477 ; (j,l) the two other axes, j < l
			const int j = a == 0 ? 1 : 0, l = a == 2 ? 1 : 2;
			double2 in = cadd(cmul(kk[j], Sl[off[a][j]]), cmul(kk[l], Sl[off[a][l]]));
			double2 t = cadd(cscale(Sl[diag[a]], g[a] - kk2), cmul(cscale(kk[a], 2.), in));
			t = cscale(t, g[a] - kk2);
			t = cadd(t, cscale(Sl[diag[j]], (g[a] + kk2) * (g[j] + kk2) - 2. * kk2 * kk2));
			t = cadd(t, cscale(Sl[diag[l]], (g[a] + kk2) * (g[l] + kk2) - 2. * kk2 * kk2));
			t = cadd(t, cmul(cmul(cscale(kk[j], 2. * (g[a] + kk2)), kk[l]), Sl[off[j][l]]));
			stc(h + diag[a] * cs + i, cdiv(t, k6));
		}
		#pragma unroll
		for (int a = 0; a < 2; a++)
			#pragma unroll
			for (int b = a + 1; b < 3; b++)
			{
				// off-diagonal (a,b): :457-459, :461-463, :470-472 ; l the third axis
				const int l = 3 - a - b;
				double2 t = cscale(Sl[off[a][b]], 2. * (g[a] - kk2) * (g[b] - kk2));
				t = cadd(t, cmul(cmul(cscale(cconj(kk[a]), g[l] + kk2), cconj(kk[b])), Sl[diag[l]]));
				double2 u = cadd(cmul(cconj(kk[a]), Sl[diag[a]]), cmul(cscale(kk[l], 2.), Sl[off[a][l]]));
				t = cadd(t, cmul(cscale(cconj(kk[b]), g[a] - kk2), u));
				double2 v = cadd(cmul(cconj(kk[b]), Sl[diag[b]]), cmul(cscale(kk[l], 2.), Sl[off[b][l]]));
				t = cadd(t, cmul(cscale(cconj(kk[a]), g[b] - kk2), v));
				stc(h + off[a][b] * cs + i, cdiv(t, k6));
			}
	}
}

int check_cplx(const gevb_field * f, int ncomp, const char * who, const char * name)
{
	GEVB_CHECK_ARG(f != NULL, "%s: %s is NULL", who, name);
	GEVB_CHECK_ARG(f->kind == GEVB_CPLX, "%s: %s must be a Fourier-space field", who, name);
	GEVB_CHECK_ARG(f->ncomp == ncomp, "%s: %s needs %d components (has %d)", who, name, ncomp, f->ncomp);
	return 0;
}

} // namespace

extern "C" int gevb_solveModifiedPoissonFT(gevb_field * sourceFT, gevb_field * potFT, double coeff, double modif)
{
	GEVB_TRY(check_cplx(sourceFT, 1, "solveModifiedPoissonFT", "sourceFT"));
	GEVB_TRY(check_cplx(potFT, 1, "solveModifiedPoissonFT", "potFT"));
	gevb_ctx * c = potFT->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_POISSON);
	KLayout L = make_klayout(c);
	coeff /= -((double) ((long) c->N * (long) c->N * (long) c->N));                                          // gevolution.hpp:511
	k_poisson<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, c->d_gridk2, (const double2 *) sourceFT->data, (double2 *) potFT->data, coeff, modif);
	KERNEL_CHECK(c);
	return 0;
}

extern "C" int gevb_projectFTscalar(gevb_field * SijFT, gevb_field * chiFT, int add)
{
	GEVB_TRY(check_cplx(SijFT, 6, "projectFTscalar", "SijFT"));
	GEVB_TRY(check_cplx(chiFT, 1, "projectFTscalar", "chiFT"));
	gevb_ctx * c = chiFT->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_FTSCALAR);
	KLayout L = make_klayout(c);
	k_ftscalar<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, c->d_gridk2, c->d_kshift, (const double2 *) SijFT->data, SijFT->comp_stride, (double2 *) chiFT->data, add);
	KERNEL_CHECK(c);
	return 0;
}

extern "C" int gevb_evolveFTvector(gevb_field * SijFT, gevb_field * BiFT, double a2dtau)
{
	GEVB_TRY(check_cplx(SijFT, 6, "evolveFTvector", "SijFT"));
	GEVB_TRY(check_cplx(BiFT, 3, "evolveFTvector", "BiFT"));
	gevb_ctx * c = BiFT->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_EVOLVE);
	KLayout L = make_klayout(c);
	k_evolve_vector<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, c->d_gridk2, c->d_kshift, (const double2 *) SijFT->data, SijFT->comp_stride, (double2 *) BiFT->data, BiFT->comp_stride, a2dtau);
	KERNEL_CHECK(c);
	return 0;
}

extern "C" int gevb_projectFTscalar_evolveFTvector(gevb_field * SijFT, gevb_field * chiFT, gevb_field * BiFT, double a2dtau)
{
	GEVB_TRY(check_cplx(SijFT, 6, "projectFTscalar_evolveFTvector", "SijFT"));
	GEVB_TRY(check_cplx(chiFT, 1, "projectFTscalar_evolveFTvector", "chiFT"));
	GEVB_TRY(check_cplx(BiFT, 3, "projectFTscalar_evolveFTvector", "BiFT"));
	GEVB_CHECK_ARG(chiFT != SijFT && BiFT != SijFT, "projectFTscalar_evolveFTvector: outputs must not alias SijFT");
	gevb_ctx * c = BiFT->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_FTSCALAR_EVOLVE);
	KLayout L = make_klayout(c);
	k_ftscalar_evolve<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, c->d_gridk2, c->d_kshift, (const double2 *) SijFT->data, SijFT->comp_stride,
		(double2 *) chiFT->data, (double2 *) BiFT->data, BiFT->comp_stride, a2dtau);
	KERNEL_CHECK(c);
	return 0;
}

extern "C" int gevb_projectFTvector(gevb_field * SiFT, gevb_field * BiFT, double coeff, double modif)
{
	GEVB_TRY(check_cplx(SiFT, 3, "projectFTvector", "SiFT"));
	GEVB_TRY(check_cplx(BiFT, 3, "projectFTvector", "BiFT"));
	gevb_ctx * c = BiFT->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_FTVECTOR);
	KLayout L = make_klayout(c);
	k_ftvector<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, c->d_gridk2, c->d_kshift, (const double2 *) SiFT->data, (double2 *) BiFT->data, BiFT->comp_stride, coeff, modif);
	KERNEL_CHECK(c);
	return 0;
}

extern "C" int gevb_projectFTtensor(gevb_field * SijFT, gevb_field * hijFT)
{
	GEVB_TRY(check_cplx(SijFT, 6, "projectFTtensor", "SijFT"));
	GEVB_TRY(check_cplx(hijFT, 6, "projectFTtensor", "hijFT"));
	gevb_ctx * c = hijFT->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_FTTENSOR);
	KLayout L = make_klayout(c);
	k_fttensor<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, c->d_gridk2, c->d_kshift, (const double2 *) SijFT->data, (double2 *) hijFT->data, hijFT->comp_stride);
	KERNEL_CHECK(c);
	return 0;
}
