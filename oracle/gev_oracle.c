/* =============================================================================
 * oracle/gev_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE
 * =============================================================================
 * Plain-C restatement of gevolution 1.2's per-step particle-mesh hot path, on
 * flat periodic arrays (no halos, no lists).  Every function cites the
 * reference file:line it follows.  It is an INDEPENDENT formulation of the same
 * algorithm (periodic index arithmetic instead of halo + fold, counting sort
 * instead of per-cell linked lists), pinned in tests/ against
 *   (1) oracle/_ref/libgevref.so -- the reference's own gevolution.hpp compiled
 *       here against the single-rank LATfield2 shim -- on seeded inputs, and
 *   (2) the analytic invariants of SURVEY.md section 4 (the reference ships no
 *       golden vectors; LATfield2-boundary semantics are "parity unpinned").
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load the resulting library.  The product never links or calls it.
 *
 * Layouts (same as oracle/ref_driver.cpp):
 *   real field    double[ncomp][N][N][N]          [c][z][y][x]
 *   Fourier field double[ncomp][N][N][N/2+1][2]   [c][kz][ky][kx][re,im]
 *   particles     pos[np][3], vel[np][3]; tensor comps (00,01,02,11,12,22)
 * Lattice dimension i of the reference = axis x,y,z for i = 0,1,2.
 * ============================================================================= */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

const char * ora_describe(void)
{
	return "plain-C restatement of gevolution 1.2 hot path (oracle/gev_oracle.c), PHINONLINEAR, GRADIENT_ORDER=1";
}

/* ---------------------------------------------------------------- helpers -- */
typedef struct { double re, im; } cplx;
static inline cplx c_make(double r, double i) { cplx z = {r, i}; return z; }
static inline cplx c_add(cplx a, cplx b) { return c_make(a.re + b.re, a.im + b.im); }
static inline cplx c_sub(cplx a, cplx b) { return c_make(a.re - b.re, a.im - b.im); }
static inline cplx c_mul(cplx a, cplx b) { return c_make(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
static inline cplx c_scale(cplx a, double s) { return c_make(a.re * s, a.im * s); }
static inline cplx c_div(cplx a, double s) { return c_make(a.re / s, a.im / s); }
static inline cplx c_conj(cplx a) { return c_make(a.re, -a.im); }

static inline int wrap(int i, int N) { return i < 0 ? i + N : (i >= N ? i - N : i); }
#define RIDX(x, y, z) (((size_t) wrap((z), N) * N + wrap((y), N)) * N + wrap((x), N))

/* gridk2[i] = (2N sin(pi i/N))^2 ; kshift[i] = 2N sin(pi i/N) e^{-i pi i/N}
 * gevolution.hpp:222-227 (identical in :296-301, :363-368, :424-429, :513-517) */
static void k_tables(int N, double * gridk2, cplx * kshift)
{
	for (int i = 0; i < N; i++)
	{
		gridk2[i] = 2. * (double) N * sin(M_PI * (double) i / (double) N);
		if (kshift) kshift[i] = c_scale(c_make(cos(M_PI * (double) i / (double) N), -sin(M_PI * (double) i / (double) N)), gridk2[i]);
		gridk2[i] *= gridk2[i];
	}
}

/* -------------------------------------------------------------------- FFT -- */
/* PlanFFT::execute (main.cpp:477,488): per-component 3-D r2c / c2r, both
 * unnormalised (manual.pdf section 4); forward kernel e^{-2 pi i k x / N}.   */
static void fft1d(int n, double * re, double * im, int sign)
{
	if ((n & (n - 1)) == 0)
	{
		for (int i = 1, j = 0; i < n; i++)
		{
			int bit = n >> 1;
			for (; j & bit; bit >>= 1) j ^= bit;
			j ^= bit;
			if (i < j) { double t = re[i]; re[i] = re[j]; re[j] = t; t = im[i]; im[i] = im[j]; im[j] = t; }
		}
		for (int len = 2; len <= n; len <<= 1)
		{
			int half = len >> 1;
			for (int k = 0; k < half; k++)
			{
				double ang = sign * 2.0 * M_PI * (double) k / (double) len;
				double c = cos(ang), s = sin(ang);
				for (int st = 0; st < n; st += len)
				{
					int a = st + k, b = a + half;
					double xr = re[b] * c - im[b] * s, xi = re[b] * s + im[b] * c;
					re[b] = re[a] - xr; im[b] = im[a] - xi; re[a] += xr; im[a] += xi;
				}
			}
		}
	}
	else
	{
		double * tr = (double *) malloc(2 * n * sizeof(double)), * ti = tr + n;
		for (int k = 0; k < n; k++)
		{
			double sr = 0., si = 0.;
			for (int j = 0; j < n; j++)
			{
				double ang = sign * 2.0 * M_PI * (double) (((long) j * k) % n) / (double) n;
				double c = cos(ang), s = sin(ang);
				sr += re[j] * c - im[j] * s; si += re[j] * s + im[j] * c;
			}
			tr[k] = sr; ti[k] = si;
		}
		memcpy(re, tr, n * sizeof(double)); memcpy(im, ti, n * sizeof(double));
		free(tr);
	}
}

void ora_fft_forward(int N, int ncomp, const double * in, double * out)
{
	const int nh = N / 2 + 1;
	const size_t V = (size_t) N * N * N, Vk = (size_t) nh * N * N;
	double * re = (double *) malloc(2 * N * sizeof(double)), * im = re + N;
	for (int c = 0; c < ncomp; c++)
	{
		const double * f = in + c * V; double * F = out + 2 * c * Vk;
		for (int z = 0; z < N; z++) for (int y = 0; y < N; y++)
		{
			for (int x = 0; x < N; x++) { re[x] = f[((size_t) z * N + y) * N + x]; im[x] = 0.; }
			fft1d(N, re, im, -1);
			for (int x = 0; x < nh; x++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); F[o] = re[x]; F[o + 1] = im[x]; }
		}
		for (int z = 0; z < N; z++) for (int x = 0; x < nh; x++)
		{
			for (int y = 0; y < N; y++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); re[y] = F[o]; im[y] = F[o + 1]; }
			fft1d(N, re, im, -1);
			for (int y = 0; y < N; y++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); F[o] = re[y]; F[o + 1] = im[y]; }
		}
		for (int y = 0; y < N; y++) for (int x = 0; x < nh; x++)
		{
			for (int z = 0; z < N; z++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); re[z] = F[o]; im[z] = F[o + 1]; }
			fft1d(N, re, im, -1);
			for (int z = 0; z < N; z++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); F[o] = re[z]; F[o + 1] = im[z]; }
		}
	}
	free(re);
}

void ora_fft_backward(int N, int ncomp, const double * in, double * out)
{
	const int nh = N / 2 + 1;
	const size_t V = (size_t) N * N * N, Vk = (size_t) nh * N * N;
	double * re = (double *) malloc(2 * N * sizeof(double)), * im = re + N;
	double * W = (double *) malloc(2 * Vk * sizeof(double));
	for (int c = 0; c < ncomp; c++)
	{
		memcpy(W, in + 2 * c * Vk, 2 * Vk * sizeof(double));
		double * f = out + c * V;
		for (int y = 0; y < N; y++) for (int x = 0; x < nh; x++)
		{
			for (int z = 0; z < N; z++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); re[z] = W[o]; im[z] = W[o + 1]; }
			fft1d(N, re, im, +1);
			for (int z = 0; z < N; z++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); W[o] = re[z]; W[o + 1] = im[z]; }
		}
		for (int z = 0; z < N; z++) for (int x = 0; x < nh; x++)
		{
			for (int y = 0; y < N; y++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); re[y] = W[o]; im[y] = W[o + 1]; }
			fft1d(N, re, im, +1);
			for (int y = 0; y < N; y++) { size_t o = 2 * (((size_t) z * N + y) * nh + x); W[o] = re[y]; W[o + 1] = im[y]; }
		}
		for (int z = 0; z < N; z++) for (int y = 0; y < N; y++)
		{
			size_t o = 2 * (((size_t) z * N + y) * nh);
			re[0] = W[o]; im[0] = 0.;
			for (int x = 1; x < nh; x++) { re[x] = W[o + 2 * x]; im[x] = W[o + 2 * x + 1]; }
			if (N % 2 == 0) im[N / 2] = 0.;
			for (int x = nh; x < N; x++) { re[x] = re[N - x]; im[x] = -im[N - x]; }
			fft1d(N, re, im, +1);
			for (int x = 0; x < N; x++) f[((size_t) z * N + y) * N + x] = re[x];
		}
	}
	free(W); free(re);
}

/* ------------------------------------------------ real-space source prep -- */
/* prepareFTsource (2), gevolution.hpp:170-192, PHINONLINEAR, not ORIGINALMETRIC */
void ora_prepareFTsource_scalar(int N, const double * phi, const double * chi, const double * source, double bgmodel, double * result, double coeff, double coeff2, double coeff3)
{
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		size_t i = RIDX(x, y, z);
		double r = coeff2 * (source[i] - bgmodel);                                 /* :176 */
		r *= 1. - 2. * phi[i];                                                     /* :184 */
		double d0 = phi[RIDX(x - 1, y, z)] - phi[RIDX(x + 1, y, z)];
		double d1 = phi[RIDX(x, y - 1, z)] - phi[RIDX(x, y + 1, z)];
		double d2 = phi[RIDX(x, y, z - 1)] - phi[RIDX(x, y, z + 1)];
		r += 0.125 * d0 * d0;                                                      /* :185 */
		r += 0.125 * d1 * d1;                                                      /* :186 */
		r += 0.125 * d2 * d2;                                                      /* :187 */
		r += (coeff3 - coeff) * phi[i] - coeff3 * chi[i];                          /* :190 */
		result[i] = r;
	}
}

/* prepareFTsource (1), gevolution.hpp:57-147, PHINONLINEAR, not ORIGINALMETRIC */
void ora_prepareFTsource_tensor(int N, const double * phi, const double * Tij, double * Sij, double coeff)
{
	const size_t V = (size_t) N * N * N;
	static const int diag[3] = {0, 3, 5};
	static const int offc[3] = {1, 2, 4};
	static const int offi[3] = {0, 0, 1}, offj[3] = {1, 2, 2};
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		size_t i = RIDX(x, y, z);
		int e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
		double p0 = phi[i];
		double pp[3], pm[3];
		for (int d = 0; d < 3; d++)
		{
			pp[d] = phi[RIDX(x + e[d][0], y + e[d][1], z + e[d][2])];
			pm[d] = phi[RIDX(x - e[d][0], y - e[d][1], z - e[d][2])];
		}
		for (int d = 0; d < 3; d++)
		{
			double s = coeff * Tij[diag[d] * V + i];                               /* :64,75,86 */
			s += 0.5 * (pp[d] - pm[d]) * (pp[d] - pm[d]);                          /* :70,81,92 */
			Sij[diag[d] * V + i] = s;
		}
		for (int k = 0; k < 3; k++)
		{
			int a = offi[k], b = offj[k];
			double pab = phi[RIDX(x + e[a][0] + e[b][0], y + e[a][1] + e[b][1], z + e[a][2] + e[b][2])];
			double s = coeff * Tij[offc[k] * V + i];                               /* :97,114,131 */
			s += pp[a] * pp[b] - p0 * pab;                                         /* :99,116,133 */
			s += 0.5 * p0 * p0;                                                    /* :106 */
			s -= 0.5 * pp[a] * pp[a];                                              /* :107 */
			s -= 0.5 * pp[b] * pp[b];                                              /* :108 */
			s += 0.5 * pab * pab;                                                  /* :109 */
			Sij[offc[k] * V + i] = s;
		}
	}
}

/* ------------------------------------------------- Fourier-space kernels -- */
#define KLOOP_BEGIN \
	const int nh = N / 2 + 1; const size_t Vk = (size_t) nh * N * N; (void) Vk; \
	for (int kz = 0; kz < N; kz++) for (int ky = 0; ky < N; ky++) for (int kx = 0; kx < nh; kx++) { \
		size_t ks = ((size_t) kz * N + ky) * nh + kx;
#define KLOOP_END }
#define CGET(F, c) c_make((F)[2 * ((c) * Vk + ks)], (F)[2 * ((c) * Vk + ks) + 1])
#define CPUT(F, c, v) do { cplx v_ = (v); (F)[2 * ((c) * Vk + ks)] = v_.re; (F)[2 * ((c) * Vk + ks) + 1] = v_.im; } while (0)

/* solveModifiedPoissonFT, gevolution.hpp:501-535 */
void ora_solveModifiedPoissonFT(int N, const double * src, double * pot, double coeff, double modif)
{
	double * gridk2 = (double *) malloc(N * sizeof(double));
	coeff /= -((long) N * (long) N * (long) N);                                    /* :511 */
	k_tables(N, gridk2, NULL);
	KLOOP_BEGIN
		cplx s = CGET(src, 0);
		if (kx == 0 && ky == 0 && kz == 0)
		{
			if (modif == 0.) CPUT(pot, 0, c_make(0., 0.));                         /* :522-523 */
			else CPUT(pot, 0, c_div(c_scale(s, coeff), modif));                    /* :525 */
		}
		else
			CPUT(pot, 0, c_div(c_scale(s, coeff), gridk2[kx] + gridk2[ky] + gridk2[kz] + modif));   /* :531 */
	KLOOP_END
	free(gridk2);
}

/* projectFTscalar, gevolution.hpp:211-265 */
void ora_projectFTscalar(int N, const double * S, double * chiFT, int add)
{
	double * g = (double *) malloc(N * sizeof(double));
	cplx * ks_ = (cplx *) malloc(N * sizeof(cplx));
	k_tables(N, g, ks_);
	KLOOP_BEGIN
		if (kx == 0 && ky == 0 && kz == 0) { CPUT(chiFT, 0, c_make(0., 0.)); continue; }      /* :230-234 */
		double g0 = g[kx], g1 = g[ky], g2 = g[kz];
		cplx k0 = ks_[kx], k1 = ks_[ky], k2 = ks_[kz];
		cplx num = c_scale(CGET(S, 0), g1 + g2 - 2. * g0);                                      /* :253 */
		num = c_add(num, c_scale(CGET(S, 3), g0 + g2 - 2. * g1));                               /* :254 */
		num = c_add(num, c_scale(CGET(S, 5), g0 + g1 - 2. * g2));                               /* :255 */
		num = c_sub(num, c_mul(c_mul(c_scale(k0, 6.), k1), CGET(S, 1)));                        /* :256 */
		num = c_sub(num, c_mul(c_mul(c_scale(k0, 6.), k2), CGET(S, 2)));                        /* :257 */
		num = c_sub(num, c_mul(c_mul(c_scale(k1, 6.), k2), CGET(S, 4)));                        /* :258 */
		cplx r = c_div(num, 2. * (g0 + g1 + g2) * (g0 + g1 + g2) * N);                          /* :259 */
		if (add) r = c_add(CGET(chiFT, 0), r);                                                  /* :240 */
		CPUT(chiFT, 0, r);
	KLOOP_END
	free(g); free(ks_);
}

/* evolveFTvector, gevolution.hpp:284-330 */
void ora_evolveFTvector(int N, const double * S, double * B, double a2dtau)
{
	double * g = (double *) malloc(N * sizeof(double));
	cplx * ks_ = (cplx *) malloc(N * sizeof(cplx));
	k_tables(N, g, ks_);
	KLOOP_BEGIN
		if (kx == 0 && ky == 0 && kz == 0) { for (int c = 0; c < 3; c++) CPUT(B, c, c_make(0., 0.)); continue; }   /* :304-310 */
		double gk[3] = {g[kx], g[ky], g[kz]};
		cplx kk[3] = {ks_[kx], ks_[ky], ks_[kz]};
		double k4 = gk[0] + gk[1] + gk[2]; k4 *= k4;                                            /* :314-315 */
		cplx Sd[3] = {CGET(S, 0), CGET(S, 3), CGET(S, 5)};
		/* off-diagonal component shared by axes (a,b): (0,1)->1, (0,2)->2, (1,2)->4 */
		static const int off[3][3] = {{-1, 1, 2}, {1, -1, 4}, {2, 4, -1}};
		cplx pref = c_make(0., -2. * a2dtau / k4);
		for (int i = 0; i < 3; i++)
		{
			int j = (i + 1) % 3, l = (i + 2) % 3;
			if (j > l) { int t = j; j = l; l = t; }                                             /* j<l as written in :317-325 */
			cplx t1 = c_scale(Sd[i], gk[j] + gk[l]);
			t1 = c_sub(t1, c_scale(Sd[j], gk[j]));
			t1 = c_sub(t1, c_scale(Sd[l], gk[l]));
			t1 = c_sub(t1, c_mul(c_mul(c_scale(kk[j], 2.), kk[l]), CGET(S, off[j][l])));
			t1 = c_mul(c_conj(kk[i]), t1);
			cplx t2 = c_add(c_mul(kk[j], CGET(S, off[i][j])), c_mul(kk[l], CGET(S, off[i][l])));
			t2 = c_scale(t2, gk[j] + gk[l] - gk[i]);
			CPUT(B, i, c_add(CGET(B, i), c_mul(pref, c_add(t1, t2))));
		}
	KLOOP_END
	free(g); free(ks_);
}

/* projectFTvector, gevolution.hpp:350-392 */
void ora_projectFTvector(int N, const double * Si, double * B, double coeff, double modif)
{
	double * g = (double *) malloc(N * sizeof(double));
	cplx * ks_ = (cplx *) malloc(N * sizeof(cplx));
	k_tables(N, g, ks_);
	KLOOP_BEGIN
		if (kx == 0 && ky == 0 && kz == 0) { for (int c = 0; c < 3; c++) CPUT(B, c, c_make(0., 0.)); continue; }   /* :371-377 */
		double k2 = g[kx] + g[ky] + g[kz];                                                      /* :381 */
		cplx kk[3] = {ks_[kx], ks_[ky], ks_[kz]};
		cplx s[3] = {CGET(Si, 0), CGET(Si, 1), CGET(Si, 2)};
		cplx tmp = c_div(c_add(c_add(c_mul(kk[0], s[0]), c_mul(kk[1], s[1])), c_mul(kk[2], s[2])), k2);   /* :383 */
		for (int i = 0; i < 3; i++)
			CPUT(B, i, c_div(c_scale(c_scale(c_sub(s[i], c_mul(c_conj(kk[i]), tmp)), 4.), coeff), k2 + modif));   /* :385-387 */
	KLOOP_END
	free(g); free(ks_);
}

/* projectFTtensor, gevolution.hpp:411-482 */
void ora_projectFTtensor(int N, const double * S, double * h)
{
	double * g = (double *) malloc(N * sizeof(double));
	cplx * ks_ = (cplx *) malloc(N * sizeof(cplx));
	k_tables(N, g, ks_);
	static const int diag[3] = {0, 3, 5};
	static const int off[3][3] = {{-1, 1, 2}, {1, -1, 4}, {2, 4, -1}};
	KLOOP_BEGIN
		if (kx == 0 && ky == 0 && kz == 0) { for (int c = 0; c < 6; c++) CPUT(h, c, c_make(0., 0.)); continue; }   /* :432-438 */
		double gk[3] = {g[kx], g[ky], g[kz]};
		cplx kk[3] = {ks_[kx], ks_[ky], ks_[kz]};
		cplx Sl[6]; for (int c = 0; c < 6; c++) Sl[c] = CGET(S, c);                             /* :442-447 */
		double k2 = gk[0] + gk[1] + gk[2];                                                      /* :449 */
		double k6 = k2 * k2 * k2 * N;                                                           /* :450 */
		for (int i = 0; i < 3; i++)
		{
			/* diagonal (i,i), :452-455, :465-468, :474-477 with (j,l) the other two axes, j<l */
			int j = (i + 1) % 3, l = (i + 2) % 3;
			if (j > l) { int t = j; j = l; l = t; }
			cplx in = c_add(c_mul(kk[j], Sl[off[i][j]]), c_mul(kk[l], Sl[off[i][l]]));
			cplx t = c_add(c_scale(Sl[diag[i]], gk[i] - k2), c_mul(c_scale(kk[i], 2.), in));
			t = c_scale(t, gk[i] - k2);
			t = c_add(t, c_scale(Sl[diag[j]], (gk[i] + k2) * (gk[j] + k2) - 2. * k2 * k2));
			t = c_add(t, c_scale(Sl[diag[l]], (gk[i] + k2) * (gk[l] + k2) - 2. * k2 * k2));
			t = c_add(t, c_mul(c_mul(c_scale(kk[j], 2. * (gk[i] + k2)), kk[l]), Sl[off[j][l]]));
			CPUT(h, diag[i], c_div(t, k6));
		}
		for (int a = 0; a < 3; a++) for (int b = a + 1; b < 3; b++)
		{
			/* off-diagonal (a,b), :457-459, :461-463, :470-472 with l the third axis */
			int l = 3 - a - b;
			cplx t = c_scale(Sl[off[a][b]], 2. * (gk[a] - k2) * (gk[b] - k2));
			t = c_add(t, c_mul(c_mul(c_scale(c_conj(kk[a]), gk[l] + k2), c_conj(kk[b])), Sl[diag[l]]));
			cplx u = c_add(c_mul(c_conj(kk[a]), Sl[diag[a]]), c_mul(c_scale(kk[l], 2.), Sl[off[a][l]]));
			t = c_add(t, c_mul(c_scale(c_conj(kk[b]), gk[a] - k2), u));
			cplx v = c_add(c_mul(c_conj(kk[b]), Sl[diag[b]]), c_mul(c_scale(kk[l], 2.), Sl[off[b][l]]));
			t = c_add(t, c_mul(c_scale(c_conj(kk[a]), gk[b] - k2), v));
			CPUT(h, off[a][b], c_div(t, k6));
		}
	KLOOP_END
	free(g); free(ks_);
}

/* ------------------------------------------------------ particle binning -- */
/* cell under which a particle is filed: floor(pos/dx) per axis, clamped
 * (LATfield2 addParticle_global / moveParticles re-filing).   */
static inline int cell_of(double p, double dx, int N)
{
	int c = (int) floor(p / dx);
	if (c >= N) c = N - 1;
	if (c < 0) c = 0;
	return c;
}

void ora_cell_index(int N, long np, const double * pos, int32_t * cell, uint32_t * counts)
{
	const double dx = 1.0 / (double) N;
	if (counts) memset(counts, 0, sizeof(uint32_t) * (size_t) N * N * N);
	for (long p = 0; p < np; p++)
	{
		int cx = cell_of(pos[3 * p], dx, N), cy = cell_of(pos[3 * p + 1], dx, N), cz = cell_of(pos[3 * p + 2], dx, N);
		int32_t key = (cz * N + cy) * N + cx;
		if (cell) cell[p] = key;
		if (counts) counts[key]++;
	}
}

/* stable counting sort of particle indices by cell key: order[] lists the
 * particles cell by cell (x fastest), in input order inside a cell -- the
 * iteration order of the reference's per-cell lists.                         */
static long * sort_by_cell(int N, long np, const double * pos, long ** start_out)
{
	const size_t V = (size_t) N * N * N;
	int32_t * cell = (int32_t *) malloc(sizeof(int32_t) * (np > 0 ? np : 1));
	long * start = (long *) calloc(V + 1, sizeof(long));
	long * order = (long *) malloc(sizeof(long) * (np > 0 ? np : 1));
	ora_cell_index(N, np, pos, cell, NULL);
	for (long p = 0; p < np; p++) start[cell[p] + 1]++;
	for (size_t c = 0; c < V; c++) start[c + 1] += start[c];
	long * fill = (long *) malloc(sizeof(long) * V);
	memcpy(fill, start, sizeof(long) * V);
	for (long p = 0; p < np; p++) order[fill[cell[p]]++] = p;
	free(fill); free(cell);
	*start_out = start;
	return order;
}

/* ------------------------------------------------------------ projections -- */
/* corner c = 4X+2Y+Z (gevolution.hpp:953) -> site offset */
#define CORNER(c, x, y, z) RIDX((x) + (((c) >> 2) & 1), (y) + (((c) >> 1) & 1), (z) + ((c) & 1))

/* projection_T00_project + projection_T00_comm, gevolution.hpp:927-1024 */
void ora_projection_T00(int N, long np, const double * pos, const double * vel, double mass_p, double a, const double * phi, double coeff, double * T00)
{
	const double dx = 1.0 / (double) N;
	const size_t V = (size_t) N * N * N;
	double mass = coeff / (dx * dx * dx); mass *= mass_p; mass /= a;              /* :945-947 */
	long * start; long * order = sort_by_cell(N, np, pos, &start);
	memset(T00, 0, sizeof(double) * V);
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		size_t cidx = ((size_t) z * N + y) * N + x;
		if (start[cidx + 1] == start[cidx]) continue;                             /* :960 */
		double ref[3] = {x * dx, y * dx, z * dx};                                 /* :962 */
		double cube[8] = {0}, cphi[8] = {0};
		if (phi) for (int c = 0; c < 8; c++) cphi[c] = phi[CORNER(c, x, y, z)];   /* :967-974 */
		for (long s = start[cidx]; s < start[cidx + 1]; s++)
		{
			long p = order[s];
			double up[3], dn[3], e = a, f = 0.;
			for (int i = 0; i < 3; i++) { up[i] = (pos[3 * p + i] - ref[i]) / dx; dn[i] = 1.0 - up[i]; }   /* :981-982 */
			if (phi)
			{
				const double * q = vel + 3 * p;
				f = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                      /* :989 */
				e = sqrt(f + a * a);                                              /* :990 */
				f = 3. * e + f / e;                                               /* :991 */
			}
			for (int c = 0; c < 8; c++)
				cube[c] += ((c & 4) ? up[0] : dn[0]) * ((c & 2) ? up[1] : dn[1]) * ((c & 1) ? up[2] : dn[2]) * (e + f * cphi[c]);   /* :995-1009 */
		}
		for (int c = 0; c < 8; c++) T00[CORNER(c, x, y, z)] += cube[c] * mass;    /* :1012-1019 (+ fold :1024) */
	}
	free(order); free(start);
}

/* scalarProjectionCIC_project + _comm (main.cpp:402,411): plain CIC, w*mass/dx^3 */
void ora_scalarProjectionCIC(int N, long np, const double * pos, double mass_p, double * rho)
{
	const double dx = 1.0 / (double) N;
	const size_t V = (size_t) N * N * N;
	const double mass = mass_p / (dx * dx * dx);
	long * start; long * order = sort_by_cell(N, np, pos, &start);
	memset(rho, 0, sizeof(double) * V);
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		size_t cidx = ((size_t) z * N + y) * N + x;
		if (start[cidx + 1] == start[cidx]) continue;
		double cube[8] = {0};
		for (long s = start[cidx]; s < start[cidx + 1]; s++)
		{
			long p = order[s];
			double up[3], dn[3];
			up[0] = (pos[3 * p] - x * dx) / dx; up[1] = (pos[3 * p + 1] - y * dx) / dx; up[2] = (pos[3 * p + 2] - z * dx) / dx;
			for (int i = 0; i < 3; i++) dn[i] = 1. - up[i];
			for (int c = 0; c < 8; c++) cube[c] += ((c & 4) ? up[0] : dn[0]) * ((c & 2) ? up[1] : dn[1]) * ((c & 1) ? up[2] : dn[2]);
		}
		for (int c = 0; c < 8; c++) rho[CORNER(c, x, y, z)] += cube[c] * mass;
	}
	free(order); free(start);
}

/* projection_T0i_project + projection_T0i_comm, gevolution.hpp:1046-1149 */
void ora_projection_T0i(int N, long np, const double * pos, const double * vel, double mass_p, const double * phi, double coeff, double * T0i)
{
	const double dx = 1.0 / (double) N;
	const size_t V = (size_t) N * N * N;
	double mass = coeff / (dx * dx * dx); mass *= mass_p;                         /* :1064-1065 */
	long * start; long * order = sort_by_cell(N, np, pos, &start);
	memset(T0i, 0, sizeof(double) * 3 * V);
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		size_t cidx = ((size_t) z * N + y) * N + x;
		if (start[cidx + 1] == start[cidx]) continue;
		double ref[3] = {x * dx, y * dx, z * dx};
		double qi[12] = {0}, cp[8] = {0};
		if (phi) for (int c = 0; c < 8; c++) cp[c] = phi[CORNER(c, x, y, z)];
		for (long s = start[cidx]; s < start[cidx + 1]; s++)
		{
			long p = order[s];
			const double * q = vel + 3 * p;
			double up[3], dn[3], w;
			for (int i = 0; i < 3; i++) { up[i] = (pos[3 * p + i] - ref[i]) / dx; dn[i] = 1.0 - up[i]; }
			w = mass * q[0];                                                      /* :1107-1112 */
			qi[0] += w * dn[1] * dn[2]; qi[1] += w * up[1] * dn[2]; qi[2] += w * dn[1] * up[2]; qi[3] += w * up[1] * up[2];
			w = mass * q[1];                                                      /* :1114-1119 */
			qi[4] += w * dn[0] * dn[2]; qi[5] += w * up[0] * dn[2]; qi[6] += w * dn[0] * up[2]; qi[7] += w * up[0] * up[2];
			w = mass * q[2];                                                      /* :1121-1126 */
			qi[8] += w * dn[0] * dn[1]; qi[9] += w * up[0] * dn[1]; qi[10] += w * dn[0] * up[1]; qi[11] += w * up[0] * up[1];
		}
		double * T0 = T0i, * T1 = T0i + V, * T2 = T0i + 2 * V;
		T0[RIDX(x, y, z)] += qi[0] * (1. + cp[0] + cp[4]);                        /* :1129 */
		T1[RIDX(x, y, z)] += qi[4] * (1. + cp[0] + cp[2]);                        /* :1130 */
		T2[RIDX(x, y, z)] += qi[8] * (1. + cp[0] + cp[1]);                        /* :1131 */
		T1[RIDX(x + 1, y, z)] += qi[5] * (1. + cp[4] + cp[6]);                    /* :1133 */
		T2[RIDX(x + 1, y, z)] += qi[9] * (1. + cp[4] + cp[5]);                    /* :1134 */
		T0[RIDX(x, y + 1, z)] += qi[1] * (1. + cp[2] + cp[6]);                    /* :1136 */
		T2[RIDX(x, y + 1, z)] += qi[10] * (1. + cp[2] + cp[3]);                   /* :1137 */
		T0[RIDX(x, y, z + 1)] += qi[2] * (1. + cp[1] + cp[5]);                    /* :1139 */
		T1[RIDX(x, y, z + 1)] += qi[6] * (1. + cp[1] + cp[3]);                    /* :1140 */
		T0[RIDX(x, y + 1, z + 1)] += qi[3] * (1. + cp[3] + cp[7]);                /* :1142 */
		T1[RIDX(x + 1, y, z + 1)] += qi[7] * (1. + cp[5] + cp[7]);                /* :1143 */
		T2[RIDX(x + 1, y + 1, z)] += qi[11] * (1. + cp[6] + cp[7]);               /* :1144 */
	}
	free(order); free(start);
}

/* projection_Tij_project + projection_Tij_comm, gevolution.hpp:1173-1300 */
void ora_projection_Tij(int N, long np, const double * pos, const double * vel, double mass_p, double a, const double * phi, double coeff, double * Tij)
{
	const double dx = 1.0 / (double) N;
	const size_t V = (size_t) N * N * N;
	double mass = coeff / (dx * dx * dx); mass *= mass_p; mass /= a;              /* :1191-1193 */
	static const int diag[3] = {0, 3, 5};
	long * start; long * order = sort_by_cell(N, np, pos, &start);
	memset(Tij, 0, sizeof(double) * 6 * V);
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		size_t cidx = ((size_t) z * N + y) * N + x;
		if (start[cidx + 1] == start[cidx]) continue;
		double ref[3] = {(double) x * dx, (double) y * dx, (double) z * dx};      /* :1210 */
		double tij[6] = {0}, tii[24] = {0}, cp[8] = {0};
		if (phi) for (int c = 0; c < 8; c++) cp[c] = phi[CORNER(c, x, y, z)];
		for (long s = start[cidx]; s < start[cidx + 1]; s++)
		{
			long p = order[s];
			const double * q = vel + 3 * p;
			double up[3], dn[3], e, f, w;
			for (int i = 0; i < 3; i++) { up[i] = (pos[3 * p + i] - ref[i]) / dx; dn[i] = 1.0 - up[i]; }
			f = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                          /* :1236 */
			e = sqrt(f + a * a);                                                  /* :1237 */
			f = 4. + a * a / (f + a * a);                                         /* :1238 */
			for (int i = 0; i < 3; i++)
			{
				w = mass * q[i] * q[i] / e;                                       /* :1243 */
				for (int c = 0; c < 8; c++)
					tii[c + i * 8] += w * ((c & 4) ? up[0] : dn[0]) * ((c & 2) ? up[1] : dn[1]) * ((c & 1) ? up[2] : dn[2]) * (1. + f * cp[c]);   /* :1245-1259 */
			}
			w = mass * q[0] * q[1] / e;                                           /* :1262-1264 */
			tij[0] += w * dn[2] * (1. + f * 0.25 * (cp[0] + cp[2] + cp[4] + cp[6]));
			tij[1] += w * up[2] * (1. + f * 0.25 * (cp[1] + cp[3] + cp[5] + cp[7]));
			w = mass * q[0] * q[2] / e;                                           /* :1266-1268 */
			tij[2] += w * dn[1] * (1. + f * 0.25 * (cp[0] + cp[1] + cp[4] + cp[5]));
			tij[3] += w * up[1] * (1. + f * 0.25 * (cp[2] + cp[3] + cp[6] + cp[7]));
			w = mass * q[1] * q[2] / e;                                           /* :1270-1272 */
			tij[4] += w * dn[0] * (1. + f * 0.25 * (cp[0] + cp[1] + cp[2] + cp[3]));
			tij[5] += w * up[0] * (1. + f * 0.25 * (cp[4] + cp[5] + cp[6] + cp[7]));
		}
		for (int i = 0; i < 3; i++) for (int c = 0; c < 8; c++) Tij[diag[i] * V + CORNER(c, x, y, z)] += tii[c + 8 * i];   /* :1277-1294 */
		Tij[1 * V + RIDX(x, y, z)] += tij[0];                                     /* :1278 */
		Tij[2 * V + RIDX(x, y, z)] += tij[2];                                     /* :1279 */
		Tij[4 * V + RIDX(x, y, z)] += tij[4];                                     /* :1280 */
		Tij[4 * V + RIDX(x + 1, y, z)] += tij[5];                                 /* :1283 */
		Tij[2 * V + RIDX(x, y + 1, z)] += tij[3];                                 /* :1286 */
		Tij[1 * V + RIDX(x, y, z + 1)] += tij[1];                                 /* :1289 */
	}
	free(order); free(start);
}

/* ------------------------------------------------------------ kick / drift -- */
/* one-sided CIC gradient of a scalar, gevolution.hpp:585-596 (GRADIENT_ORDER 1) */
static void grad_cic(int N, const double * f, int x, int y, int z, const double * r, double * g, double sign)
{
#define P(dx_, dy_, dz_) f[RIDX(x + (dx_), y + (dy_), z + (dz_))]
	double g0, g1, g2;
	g0 = (1. - r[1]) * (1. - r[2]) * (P(1, 0, 0) - P(0, 0, 0));
	g1 = (1. - r[0]) * (1. - r[2]) * (P(0, 1, 0) - P(0, 0, 0));
	g2 = (1. - r[0]) * (1. - r[1]) * (P(0, 0, 1) - P(0, 0, 0));
	g0 += r[1] * (1. - r[2]) * (P(1, 1, 0) - P(0, 1, 0));
	g1 += r[0] * (1. - r[2]) * (P(1, 1, 0) - P(1, 0, 0));
	g2 += r[0] * (1. - r[1]) * (P(1, 0, 1) - P(1, 0, 0));
	g0 += (1. - r[1]) * r[2] * (P(1, 0, 1) - P(0, 0, 1));
	g1 += (1. - r[0]) * r[2] * (P(0, 1, 1) - P(0, 0, 1));
	g2 += (1. - r[0]) * r[1] * (P(0, 1, 1) - P(0, 1, 0));
	g0 += r[1] * r[2] * (P(1, 1, 1) - P(0, 1, 1));
	g1 += r[0] * r[2] * (P(1, 1, 1) - P(1, 0, 1));
	g2 += r[0] * r[1] * (P(1, 1, 1) - P(1, 1, 0));
#undef P
	if (sign > 0) { g[0] = g0; g[1] = g1; g[2] = g2; }
	else { g[0] -= g0; g[1] -= g1; g[2] -= g2; }
}

/* Particles::updateVel driver + update_q (gevolution.hpp:570-678)
 * / update_q_Newton (:709-776).  Returns sqrt(max v2).                        */
double ora_updateVel(int N, long np, const double * pos, double * vel, int kind, double dtau, const double * phi, const double * chi, const double * Bi, int nfields, const double * params)
{
	const double dx = 1.0 / (double) N;
	const size_t V = (size_t) N * N * N;
	double maxv2 = 0.;
	for (long p = 0; p < np; p++)
	{
		double r[3], ip;
		int x = cell_of(pos[3 * p], dx, N), y = cell_of(pos[3 * p + 1], dx, N), z = cell_of(pos[3 * p + 2], dx, N);
		for (int l = 0; l < 3; l++) r[l] = modf(pos[3 * p + l] / dx, &ip);
		double * q = vel + 3 * p;
		double g[3], v2;
		if (kind == 0)
		{
			double pg[3] = {0, 0, 0};
			v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                         /* :581 */
			double e2 = v2 + params[0] * params[0];                               /* :582 */
			grad_cic(N, phi, x, y, z, r, g, +1.);                                 /* :585-596 */
			g[0] *= (v2 + e2) / e2; g[1] *= (v2 + e2) / e2; g[2] *= (v2 + e2) / e2;   /* :613-615 */
			if (nfields >= 2 && chi != NULL) grad_cic(N, chi, x, y, z, r, g, -1.);    /* :617-631 */
			e2 = sqrt(e2);                                                        /* :633 */
			if (nfields >= 3 && Bi != NULL)
			{
#define B(c, dx_, dy_, dz_) Bi[(c) * V + RIDX(x + (dx_), y + (dy_), z + (dz_))]
				/* :637-642 */
				pg[0] = ((1. - r[2]) * (B(1, 1, 0, 0) - B(1, 0, 0, 0)) + r[2] * (B(1, 1, 0, 1) - B(1, 0, 0, 1))) * q[1];
				pg[0] += ((1. - r[1]) * (B(2, 1, 0, 0) - B(2, 0, 0, 0)) + r[1] * (B(2, 1, 1, 0) - B(2, 0, 1, 0))) * q[2];
				pg[0] += (1. - r[1]) * (1. - r[2]) * ((r[0] - 1.) * B(0, -1, 0, 0) + (1. - 2. * r[0]) * B(0, 0, 0, 0) + r[0] * B(0, 1, 0, 0)) * q[0];
				pg[0] += r[1] * (1. - r[2]) * ((r[0] - 1.) * B(0, -1, 1, 0) + (1. - 2. * r[0]) * B(0, 0, 1, 0) + r[0] * B(0, 1, 1, 0)) * q[0];
				pg[0] += (1. - r[1]) * r[2] * ((r[0] - 1.) * B(0, -1, 0, 1) + (1. - 2. * r[0]) * B(0, 0, 0, 1) + r[0] * B(0, 1, 0, 1)) * q[0];
				pg[0] += r[1] * r[2] * ((r[0] - 1.) * B(0, -1, 1, 1) + (1. - 2. * r[0]) * B(0, 0, 1, 1) + r[0] * B(0, 1, 1, 1)) * q[0];
				/* :644-649 */
				pg[1] = ((1. - r[0]) * (B(2, 0, 1, 0) - B(2, 0, 0, 0)) + r[0] * (B(2, 1, 1, 0) - B(2, 1, 0, 0))) * q[2];
				pg[1] += ((1. - r[2]) * (B(0, 0, 1, 0) - B(0, 0, 0, 0)) + r[2] * (B(0, 0, 1, 1) - B(0, 0, 0, 1))) * q[0];
				pg[1] += (1. - r[0]) * (1. - r[2]) * ((r[1] - 1.) * B(1, 0, -1, 0) + (1. - 2. * r[1]) * B(1, 0, 0, 0) + r[1] * B(1, 0, 1, 0)) * q[1];
				pg[1] += r[0] * (1. - r[2]) * ((r[1] - 1.) * B(1, 1, -1, 0) + (1. - 2. * r[1]) * B(1, 1, 0, 0) + r[1] * B(1, 1, 1, 0)) * q[1];
				pg[1] += (1. - r[0]) * r[2] * ((r[1] - 1.) * B(1, 0, -1, 1) + (1. - 2. * r[1]) * B(1, 0, 0, 1) + r[1] * B(1, 0, 1, 1)) * q[1];
				pg[1] += r[0] * r[2] * ((r[1] - 1.) * B(1, 1, -1, 1) + (1. - 2. * r[1]) * B(1, 1, 0, 1) + r[1] * B(1, 1, 1, 1)) * q[1];
				/* :651-656 */
				pg[2] = ((1. - r[1]) * (B(0, 0, 0, 1) - B(0, 0, 0, 0)) + r[1] * (B(0, 0, 1, 1) - B(0, 0, 1, 0))) * q[0];
				pg[2] += ((1. - r[0]) * (B(1, 0, 0, 1) - B(1, 0, 0, 0)) + r[0] * (B(1, 1, 0, 1) - B(1, 1, 0, 0))) * q[1];
				pg[2] += (1. - r[0]) * (1. - r[1]) * ((r[2] - 1.) * B(2, 0, 0, -1) + (1. - 2. * r[2]) * B(2, 0, 0, 0) + r[2] * B(2, 0, 0, 1)) * q[2];
				pg[2] += r[0] * (1. - r[1]) * ((r[2] - 1.) * B(2, 1, 0, -1) + (1. - 2. * r[2]) * B(2, 1, 0, 0) + r[2] * B(2, 1, 0, 1)) * q[2];
				pg[2] += (1. - r[0]) * r[1] * ((r[2] - 1.) * B(2, 0, 1, -1) + (1. - 2. * r[2]) * B(2, 0, 1, 0) + r[2] * B(2, 0, 1, 1)) * q[2];
				pg[2] += r[0] * r[1] * ((r[2] - 1.) * B(2, 1, 1, -1) + (1. - 2. * r[2]) * B(2, 1, 1, 0) + r[2] * B(2, 1, 1, 1)) * q[2];
#undef B
				g[0] += pg[0] / params[1] / e2;                                   /* :658-660 */
				g[1] += pg[1] / params[1] / e2;
				g[2] += pg[2] / params[1] / e2;
			}
			v2 = 0.;
			for (int i = 0; i < 3; i++) { q[i] -= dtau * e2 * g[i] / dx; v2 += q[i] * q[i]; }   /* :664-668 */
			v2 = v2 / params[0] / params[0];                                      /* :670 */
		}
		else
		{
			grad_cic(N, phi, x, y, z, r, g, +1.);                                 /* :719-730 */
			if (nfields >= 2 && chi != NULL) grad_cic(N, chi, x, y, z, r, g, -1.);    /* :747-761 */
			v2 = 0.;
			for (int i = 0; i < 3; i++) { q[i] -= dtau * params[0] * g[i] / dx; v2 += q[i] * q[i]; }   /* :764-768 */
			v2 = v2 / params[0] / params[0];                                      /* :770 */
		}
		if (v2 > maxv2) maxv2 = v2;
	}
	return sqrt(maxv2);
}

/* periodic wrap convention (edge semantics (1)); box length 1 */
static inline double wrap_pos(double p)
{
	double w = p - floor(p / 1.0) * 1.0;
	if (w >= 1.0) w = 0.;
	return w;
}

/* Particles::moveParticles driver + update_pos (gevolution.hpp:810-871) /
 * update_pos_Newton (:900-903); positions come back wrapped into [0,1).       */
void ora_moveParticles(int N, long np, double * pos, const double * vel, int kind, double dtau, const double * phi, const double * chi, const double * Bi, int nfields, const double * params)
{
	const double dx = 1.0 / (double) N;
	const size_t V = (size_t) N * N * N;
	for (long p = 0; p < np; p++)
	{
		double * xp = pos + 3 * p; const double * q = vel + 3 * p;
		if (kind != 0)
		{
			for (int l = 0; l < 3; l++) xp[l] += dtau * q[l] / params[0];         /* :902 */
		}
		else
		{
			double r[3], ip, v[3];
			int x = cell_of(xp[0], dx, N), y = cell_of(xp[1], dx, N), z = cell_of(xp[2], dx, N);
			for (int l = 0; l < 3; l++) r[l] = modf(xp[l] / dx, &ip);
			double v2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];                  /* :813 */
			double e2 = v2 + params[0] * params[0];                               /* :814 */
			double ph = 0., ch = 0.;
			if (nfields >= 1)
			{                                                                     /* :820-827 */
				ph = phi[RIDX(x, y, z)] * (1. - r[0]) * (1. - r[1]) * (1. - r[2]);
				ph += phi[RIDX(x + 1, y, z)] * r[0] * (1. - r[1]) * (1. - r[2]);
				ph += phi[RIDX(x, y + 1, z)] * (1. - r[0]) * r[1] * (1. - r[2]);
				ph += phi[RIDX(x + 1, y + 1, z)] * r[0] * r[1] * (1. - r[2]);
				ph += phi[RIDX(x, y, z + 1)] * (1. - r[0]) * (1. - r[1]) * r[2];
				ph += phi[RIDX(x + 1, y, z + 1)] * r[0] * (1. - r[1]) * r[2];
				ph += phi[RIDX(x, y + 1, z + 1)] * (1. - r[0]) * r[1] * r[2];
				ph += phi[RIDX(x + 1, y + 1, z + 1)] * r[0] * r[1] * r[2];
			}
			if (nfields >= 2)
			{                                                                     /* :832-839 */
				ch = chi[RIDX(x, y, z)] * (1. - r[0]) * (1. - r[1]) * (1. - r[2]);
				ch += chi[RIDX(x + 1, y, z)] * r[0] * (1. - r[1]) * (1. - r[2]);
				ch += chi[RIDX(x, y + 1, z)] * (1. - r[0]) * r[1] * (1. - r[2]);
				ch += chi[RIDX(x + 1, y + 1, z)] * r[0] * r[1] * (1. - r[2]);
				ch += chi[RIDX(x, y, z + 1)] * (1. - r[0]) * (1. - r[1]) * r[2];
				ch += chi[RIDX(x + 1, y, z + 1)] * r[0] * (1. - r[1]) * r[2];
				ch += chi[RIDX(x, y + 1, z + 1)] * (1. - r[0]) * r[1] * r[2];
				ch += chi[RIDX(x + 1, y + 1, z + 1)] * r[0] * r[1] * r[2];
			}
			v2 = (1. + (3. - v2 / e2) * ph - ch) / sqrt(e2);                      /* :842 */
			v[0] = q[0] * v2; v[1] = q[1] * v2; v[2] = q[2] * v2;                 /* :844-846 */
			if (nfields >= 3)
			{
				double b[3];
#define B(c, dx_, dy_, dz_) Bi[(c) * V + RIDX(x + (dx_), y + (dy_), z + (dz_))]
				b[0] = B(0, 0, 0, 0) * (1. - r[1]) * (1. - r[2]);                 /* :852 */
				b[1] = B(1, 0, 0, 0) * (1. - r[0]) * (1. - r[2]);                 /* :853 */
				b[2] = B(2, 0, 0, 0) * (1. - r[0]) * (1. - r[1]);                 /* :854 */
				b[1] += B(1, 1, 0, 0) * r[0] * (1. - r[2]);                       /* :855 */
				b[2] += B(2, 1, 0, 0) * r[0] * (1. - r[1]);                       /* :856 */
				b[0] += B(0, 0, 1, 0) * r[1] * (1. - r[2]);                       /* :857 */
				b[2] += B(2, 0, 1, 0) * (1. - r[0]) * r[1];                       /* :858 */
				b[0] += B(0, 0, 0, 1) * (1. - r[1]) * r[2];                       /* :859 */
				b[1] += B(1, 0, 0, 1) * (1. - r[0]) * r[2];                       /* :860 */
				b[1] += B(1, 1, 0, 1) * r[0] * r[2];                              /* :861 */
				b[0] += B(0, 0, 1, 1) * r[1] * r[2];                              /* :862 */
				b[2] += B(2, 1, 1, 0) * r[0] * r[1];                              /* :863 */
#undef B
				for (int l = 0; l < 3; l++) xp[l] += dtau * (v[l] + b[l] / params[1]);   /* :865 */
			}
			else
				for (int l = 0; l < 3; l++) xp[l] += dtau * v[l];                 /* :869 */
		}
		for (int l = 0; l < 3; l++) xp[l] = wrap_pos(xp[l]);
	}
}

/* ---------------------------------------------------------------- analysis -- */
/* extractCrossSpectrum with fld1 == fld2, tools.hpp:53-212 */
void ora_extractPowerSpectrum(int N, int ncomp, int symm, const double * F, double * kbin, double * power, double * kscatter, double * pscatter, int * occupation, int numbins, int deconvolve, int ktype)
{
	double * typek2 = (double *) malloc(N * sizeof(double)), * sinc = (double *) malloc(N * sizeof(double));
	int i;
	if (ktype == 0) { for (i = 0; i < N; i++) { typek2[i] = 2. * (double) N * sin(M_PI * (double) i / (double) N); typek2[i] *= typek2[i]; } }   /* :66-73 */
	else
	{
		for (i = 0; i <= N / 2; i++) { typek2[i] = 2. * M_PI * (double) i; typek2[i] *= typek2[i]; }          /* :76-80 */
		for (; i < N; i++) { typek2[i] = 2. * M_PI * (double) (N - i); typek2[i] *= typek2[i]; }              /* :81-85 */
	}
	sinc[0] = 1.;
	for (i = 1; i <= N / 2; i++) sinc[i] = deconvolve ? sin(M_PI * (float) i / (float) N) * (float) N / (M_PI * (float) i) : 1.;   /* :91-93 (float casts as written) */
	for (; i < N; i++) sinc[i] = sinc[N - i];                                                                 /* :103-106 */
	double k2max = 3. * typek2[N / 2];                                                                        /* :108 */
	for (i = 0; i < numbins; i++) { kbin[i] = power[i] = kscatter[i] = pscatter[i] = 0.; occupation[i] = 0; }
	KLOOP_BEGIN
		int weight;
		if (kx == 0 && ky == 0 && kz == 0) continue;                                                          /* :121-122 */
		else if (kx == 0) weight = 1;
		else if ((kx == N / 2) && (N % 2 == 0)) weight = 1;
		else weight = 2;
		double k2 = typek2[kx] + typek2[ky] + typek2[kz];                                                     /* :130 */
		double s = sinc[kx] * sinc[ky] * sinc[kz]; s *= s;                                                    /* :131-132 */
		double p;
		if (symm)
		{                                                                                                     /* :138-147 */
			cplx a1 = CGET(F, 1), a2 = CGET(F, 2), a4 = CGET(F, 4), a0 = CGET(F, 0), a3 = CGET(F, 3), a5 = CGET(F, 5);
			p = a1.re * a1.re + a1.im * a1.im; p += a2.re * a2.re + a2.im * a2.im; p += a4.re * a4.re + a4.im * a4.im;
			p *= 2.;
			p += a0.re * a0.re + a0.im * a0.im; p += a3.re * a3.re + a3.im * a3.im; p += a5.re * a5.re + a5.im * a5.im;
		}
		else
		{                                                                                                     /* :149-153 */
			p = 0.;
			for (int c = 0; c < ncomp; c++) { cplx v = CGET(F, c); p += v.re * v.re + v.im * v.im; }
		}
		i = (int) floor((double) ((double) numbins * sqrt(k2 / k2max)));                                      /* :155 */
		if (i < numbins)
		{                                                                                                     /* :158-162 */
			kbin[i] += weight * sqrt(k2);
			kscatter[i] += weight * k2;
			power[i] += weight * p * k2 * sqrt(k2) / s;
			pscatter[i] += weight * p * p * k2 * k2 * k2 / s / s;
			occupation[i] += weight;
		}
	KLOOP_END
	for (i = 0; i < numbins; i++)
	{
		if (occupation[i] > 0)
		{                                                                                                     /* :186-193 */
			kscatter[i] = sqrt(kscatter[i] * occupation[i] - kbin[i] * kbin[i]) / occupation[i];
			if (!isfinite(kscatter[i])) kscatter[i] = 0.;
			kbin[i] = kbin[i] / occupation[i];
			power[i] /= occupation[i];
			pscatter[i] = sqrt(pscatter[i] / occupation[i] - power[i] * power[i]);
			if (!isfinite(pscatter[i])) pscatter[i] = 0.;
		}
	}
	free(typek2); free(sinc);
}

/* computeVectorDiagnostics, tools.hpp:364-386 */
void ora_computeVectorDiagnostics(int N, const double * Bi, double * mdivB, double * mcurlB)
{
	const size_t V = (size_t) N * N * N;
	const double ls = (double) N;
	double md = 0., mc = 0.;
#define B(c, dx_, dy_, dz_) Bi[(c) * V + RIDX(x + (dx_), y + (dy_), z + (dz_))]
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		double b1 = fabs((B(0, 0, 0, 0) - B(0, -1, 0, 0)) + (B(1, 0, 0, 0) - B(1, 0, -1, 0)) + (B(2, 0, 0, 0) - B(2, 0, 0, -1))) * ls;
		if (b1 > md) md = b1;
		b1 = 0.5 * (B(0, 0, 0, 0) + B(1, 1, 0, 0) - B(0, 0, 1, 0) - B(1, 0, 0, 0) + B(0, 0, 0, 1) + B(1, 1, 0, 1) - B(0, 0, 1, 1) - B(1, 0, 0, 1)) * ls;
		double b2 = 0.5 * (B(0, 0, 0, 0) + B(2, 1, 0, 0) - B(0, 0, 0, 1) - B(2, 0, 0, 0) + B(0, 0, 1, 0) + B(2, 1, 1, 0) - B(0, 0, 1, 1) - B(2, 0, 1, 0)) * ls;
		double b3 = 0.5 * (B(2, 0, 0, 0) + B(1, 0, 0, 1) - B(2, 0, 1, 0) - B(1, 0, 0, 0) + B(2, 1, 0, 0) + B(1, 1, 0, 1) - B(2, 1, 1, 0) - B(1, 1, 0, 0)) * ls;
		double b4 = sqrt(b1 * b1 + b2 * b2 + b3 * b3);
		if (b4 > mc) mc = b4;
	}
#undef B
	*mdivB = md; *mcurlB = mc;
}

/* computeTensorDiagnostics, tools.hpp:405-431 */
void ora_computeTensorDiagnostics(int N, const double * hij, double * mdivh, double * mtraceh, double * mnormh)
{
	const size_t V = (size_t) N * N * N;
	const double ls = (double) N;
	double md = 0., mt = 0., mn = 0.;
#define H(c, dx_, dy_, dz_) hij[(c) * V + RIDX(x + (dx_), y + (dy_), z + (dz_))]
	for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++)
	{
		double d1 = (H(0, 1, 0, 0) - H(0, 0, 0, 0) + H(1, 0, 0, 0) - H(1, 0, -1, 0) + H(2, 0, 0, 0) - H(2, 0, 0, -1)) * ls;
		double d2 = (H(3, 0, 1, 0) - H(3, 0, 0, 0) + H(1, 0, 0, 0) - H(1, -1, 0, 0) + H(4, 0, 0, 0) - H(4, 0, 0, -1)) * ls;
		double d3 = (H(5, 0, 0, 1) - H(5, 0, 0, 0) + H(2, 0, 0, 0) - H(2, -1, 0, 0) + H(4, 0, 0, 0) - H(4, 0, -1, 0)) * ls;
		d1 = sqrt(d1 * d1 + d2 * d2 + d3 * d3);
		if (d1 > md) md = d1;
		d1 = fabs(H(0, 0, 0, 0) + H(3, 0, 0, 0) + H(5, 0, 0, 0));
		if (d1 > mt) mt = d1;
		d1 = sqrt(H(0, 0, 0, 0) * H(0, 0, 0, 0) + 2. * H(1, 0, 0, 0) * H(1, 0, 0, 0) + 2. * H(2, 0, 0, 0) * H(2, 0, 0, 0) + H(3, 0, 0, 0) * H(3, 0, 0, 0) + 2. * H(4, 0, 0, 0) * H(4, 0, 0, 0) + H(5, 0, 0, 0) * H(5, 0, 0, 0));
		if (d1 > mn) mn = d1;
	}
#undef H
	*mdivh = md; *mtraceh = mt; *mnormh = mn;
}

/* -------------------------------------------------------------- background -- */
/* cosmo: Omega_cdm, Omega_b, Omega_m, Omega_Lambda, Omega_fld, w0_fld, wa_fld, Omega_g, Omega_ur, Omega_rad, h
 * (no ncdm species: bg_ncdm == 0).  Hconf: background.hpp:137-140             */
double ora_Hconf(double a, double fourpiG, const double * c)
{
	return sqrt((2. * fourpiG / 3.) * (((c[0] + c[1] + 0.) / a) + (c[3] * a * a) + (c[9] / a / a) + (c[4] * exp(3. * c[6] * (a - 1.)) / pow(a, 1. + 3. * (c[5] + c[6])))));
}

/* rungekutta4bg: background.hpp:167-177 */
double ora_rungekutta4bg(double a, double fourpiG, const double * c, double dtau)
{
	double k1a = a * ora_Hconf(a, fourpiG, c);
	double k2a = (a + k1a * dtau / 2.) * ora_Hconf(a + k1a * dtau / 2., fourpiG, c);
	double k3a = (a + k2a * dtau / 2.) * ora_Hconf(a + k2a * dtau / 2., fourpiG, c);
	double k4a = (a + k3a * dtau) * ora_Hconf(a + k3a * dtau, fourpiG, c);
	return a + dtau * (k1a + 2. * k2a + 2. * k3a + k4a) / 6.;
}
