// xpass.cu -- own x-pass of the forward 3-D transform (real to complex along x) with the source preparation fused into its
//             load:  prepareFTsource (gevolution.hpp:170-192 scalar, :57-147 tensor) + the first of the three passes of
//             PlanFFT::execute(FFT_FORWARD) in ONE pass over HBM.
//
// Separately, prepareFTsource reads and writes every component once (2 x 8 B per site) and cuFFT's x-pass reads it again
// and writes the half spectrum (2 x 8 B per site): 32 B per site and component.  Fused, the prepared values never travel:
// 16 B per site and component (the phi stencil comes from L1 / L2).  cuFFT cannot run the y-pass alone on the [z][y][kx]
// layout in one call (as a strided 2-D (z, y) transform its passes take 1.1 ms instead of 0.43 ms), so the y-pass is an own
// kernel too (k_ypass: eight kx columns per block, two 256-point FFTs + a radix-2 step per column); the z-pass is cuFFT's.
//
// STATUS: correct (1e-12 against cuFFT and numpy at 512^3), traffic ideal, but both kernels are latency-bound (no load /
// compute overlap inside a warp or block): x-pass 0.85 ms (scalar, with preparation) and 0.95 ms per tensor component, y-pass
// 0.87 ms, against cuFFT's 0.52 and 0.43 ms + 0.45 ms of separate preparation.  Off by default (knob fft_xpass); DESIGN.md 5.2.
//
// One warp transforms one row of N = 512 reals.  The row is read as 256 complex numbers z[n] = x[2n] + i x[2n+1] (which is
// how it lies in memory), transformed by a 256-point complex FFT and un-packed into the 257 coefficients of the real
// transform.  256 = 8 x 8 x 4: lane b holds z[32 a + b] (eight coalesced 16-byte loads),
//     stage 1   radix 8 over a            -> A[b][c],  times w256^(b c)
//     stage 2   radix 8 over e, b = 4e+f  -> B[c][f][g], times w32^(f g)
//     stage 3   radix 4 over f            -> Z[c + 8 g + 64 h]
// with two exchanges through a private 4.5 KB of shared memory per warp (only __syncwarp, no block barrier; layouts padded
// so that every 16-byte access is free of bank conflicts).  After stage 3 lane l holds Z[l + 32 m], m = 0..7; the partner
// Z[256 - k] of the un-packing comes through one more exchange, and the lane stores X[l + 32 m] (512 contiguous bytes per
// instruction).  Twiddle factors come from tables computed on the host in double precision.
//
// Only N = 512 on a single rank (the size the metric is quoted on); every other case keeps the cuFFT path.  The two paths are
// compared with each other and with numpy at full size in tests/test_gpu_parity.py::test_own_xpass_512.
#include "gevb_internal.cuh"
#include <math.h>

namespace {

#define XP_N 512
#define XP_M 256                          // complex points of a packed row
#define XP_NH 257
#define XP_WS 288                         // double2 per warp of exchange space: max(8 x 36, 4 x 66, 256)
#define XP_TAB (256 + 36 + 256)           // w256^(b c) as [c][b]; w32^(f g) as [f][9]; w512^k
#define XP_WARPS_PLAIN 8
#define XP_WARPS_TENSOR 6                 // one row x six components

struct XParams
{
	int nzl;                              // planes
	const double * src;                   // MODE 0 / 1: the real component (bulk, without the lower ghost plane); MODE 2: T (6 components)
	size_t cs_src;
	const double * phi, * chi;            // ghosted fields, MODE 1 / 2
	double2 * out;                        // [ncomp][nzl][N][257]
	size_t cs_out;
	const double2 * tab;
	double bgmodel, coeff, coeff2, coeff3;
	double * partial;                     // MODE 1: per-row sums of the incoming source, or NULL
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 mul_mi(double2 a) { return make_double2(a.y, -a.x); }      // a * (-i)

// forward DFT of four points, in place (natural order)
__device__ __forceinline__ void dft4(double2 & x0, double2 & x1, double2 & x2, double2 & x3)
{
	const double2 t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = mul_mi(csub(x1, x3));
	x0 = cadd(t0, t2); x2 = csub(t0, t2); x1 = cadd(t1, t3); x3 = csub(t1, t3);
}

// forward DFT of eight points, in place (natural order): two DFT-4 of the even / odd inputs and the w8^k butterflies
__device__ __forceinline__ void dft8(double2 * a)
{
	double2 e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7];
	dft4(e0, e1, e2, e3);
	dft4(o0, o1, o2, o3);
	const double r = 0.70710678118654752440;
	const double2 w1 = make_double2(r * (o1.x + o1.y), r * (o1.y - o1.x));           // o1 * (1 - i) / sqrt 2
	const double2 w2 = mul_mi(o2);
	const double2 w3 = make_double2(r * (o3.y - o3.x), -r * (o3.x + o3.y));          // o3 * (-1 - i) / sqrt 2
	a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
	a[1] = cadd(e1, w1); a[5] = csub(e1, w1);
	a[2] = cadd(e2, w2); a[6] = csub(e2, w2);
	a[3] = cadd(e3, w3); a[7] = csub(e3, w3);
}

// 256-point forward FFT of a warp: z[m] = element lane + 32 m on entry, Z[lane + 32 m] on exit; ws = the warp's exchange space
__device__ __forceinline__ void fft256(double2 * z, double2 * ws, const double2 * tab, int lane)
{
	const double2 * T1 = tab, * T2 = tab + 256;
	// ---- stage 1
	dft8(z);
	#pragma unroll
	for (int c = 1; c < 8; c++) z[c] = cmul(z[c], T1[c * 32 + lane]);
	#pragma unroll
	for (int c = 0; c < 8; c++) ws[c * 36 + lane] = z[c];
	__syncwarp();
	const int c2 = lane >> 2, f2 = lane & 3;
	#pragma unroll
	for (int e = 0; e < 8; e++) z[e] = ws[c2 * 36 + 4 * e + f2];
	__syncwarp();
	// ---- stage 2
	dft8(z);
	#pragma unroll
	for (int g = 1; g < 8; g++) z[g] = cmul(z[g], T2[f2 * 9 + g]);
	#pragma unroll
	for (int g = 0; g < 8; g++) ws[f2 * 66 + g * 8 + c2] = z[g];
	__syncwarp();
	double2 u[2][4];
	#pragma unroll
	for (int j = 0; j < 2; j++)
		#pragma unroll
		for (int f = 0; f < 4; f++) u[j][f] = ws[f * 66 + lane + 32 * j];
	__syncwarp();
	// ---- stage 3: Z[lane + 32 (j + 2 h)]
	#pragma unroll
	for (int j = 0; j < 2; j++)
	{
		dft4(u[j][0], u[j][1], u[j][2], u[j][3]);
		#pragma unroll
		for (int h = 0; h < 4; h++) z[j + 2 * h] = u[j][h];
	}
}

// z[m] = packed row element lane + 32 m on entry; X[k] for k = lane + 32 m (and X[256] by lane 0) stored to `out` on exit
__device__ __forceinline__ void row_fft_store(double2 * z, double2 * ws, const double2 * tab, double2 * out, int lane)
{
	const double2 * T3 = tab + 256 + 36;
	fft256(z, ws, tab, lane);
	// ---- un-packing: X[k] = (Z[k] + conj Z[256-k]) / 2 - i w512^k (Z[k] - conj Z[256-k]) / 2
	#pragma unroll
	for (int m = 0; m < 8; m++) ws[m * 32 + lane] = z[m];
	__syncwarp();
	#pragma unroll
	for (int m = 0; m < 8; m++)
	{
		const int k = lane + 32 * m;
		const double2 zc = ws[(XP_M - k) & (XP_M - 1)], w = T3[k];
		const double sx = z[m].x + zc.x, sy = z[m].y - zc.y, dx = z[m].x - zc.x, dy = z[m].y + zc.y;
		out[k] = make_double2(0.5 * (sx + (w.x * dy + w.y * dx)), 0.5 * (sy - (w.x * dx - w.y * dy)));
	}
	if (lane == 0) out[XP_M] = make_double2(z[0].x - z[0].y, 0.);
	__syncwarp();
}

__device__ __forceinline__ double2 ld2(const double * p) { return *(const double2 *) p; }
__device__ __forceinline__ double2 ld2s(const double * p) { return __ldcs((const double2 *) p); }

// gevolution.hpp:176-190 for one site (the same operations in the same order as k_prepare_scalar)
__device__ __forceinline__ double scalar_site(double src, double p, double chi, double pxm, double pxp, double pym, double pyp, double pzm, double pzp,
                                              double bgmodel, double coeff, double coeff2, double coeff3)
{
	double res = coeff2 * (src - bgmodel);
	res *= 1. - 2. * p;
	const double d0 = pxm - pxp, d1 = pym - pyp, d2 = pzm - pzp;
	res += 0.125 * d0 * d0;
	res += 0.125 * d1 * d1;
	res += 0.125 * d2 * d2;
	res += (coeff3 - coeff) * p - coeff3 * chi;
	return res;
}

// MODE 0: transform src as it is; MODE 1: scalar prepareFTsource on load; MODE 2: tensor prepareFTsource on load, one warp per
// (row, component) -- the six warps of a row read the same phi stencil at the same time, so L1 serves five of the six
template <int MODE>
__global__ void __launch_bounds__(MODE == 2 ? XP_WARPS_TENSOR * 32 : XP_WARPS_PLAIN * 32, MODE == 2 ? 5 : 4) k_xpass(XParams P)
{
	extern __shared__ __align__(16) double2 xp_smem[];
	constexpr int WARPS = MODE == 2 ? XP_WARPS_TENSOR : XP_WARPS_PLAIN;
	double2 * tab = xp_smem;
	for (int i = threadIdx.x; i < XP_TAB; i += WARPS * 32) tab[i] = P.tab[i];
	__syncthreads();
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double2 * ws = xp_smem + XP_TAB + warp * XP_WS;
	const int N = XP_N;
	const size_t plane = (size_t) N * N;
	const int rows = P.nzl * N;
	const int comp = MODE == 2 ? warp % 6 : 0;
	const int rows_per_block = MODE == 2 ? WARPS / 6 : WARPS;
	for (int row = blockIdx.x * rows_per_block + (MODE == 2 ? warp / 6 : warp); row < rows; row += gridDim.x * rows_per_block)
	{
		const int zl = row / N, y = row - zl * N;
		double2 z[8];
		if (MODE == 0)
		{
			const double2 * r = (const double2 *) (P.src + (size_t) row * N);
			#pragma unroll
			for (int m = 0; m < 8; m++) z[m] = __ldcs(r + lane + 32 * m);
		}
		else
		{
			const size_t grow = ((size_t) (zl + 1) * N + y) * N;                    // this row in a ghosted field
			const size_t growm = ((size_t) (zl + 1) * N + (y == 0 ? N - 1 : y - 1)) * N, growp = ((size_t) (zl + 1) * N + (y == N - 1 ? 0 : y + 1)) * N;
			double acc = 0.;
			#pragma unroll
			for (int m = 0; m < 8; m++)
			{
				const int x = 2 * (lane + 32 * m);
				const int xm = x == 0 ? N - 1 : x - 1, xp = x == N - 2 ? 0 : x + 2;
				const size_t s = grow + x;
				if (MODE == 1)
				{
					const double2 p = ld2(P.phi + s), ym = ld2(P.phi + growm + x), yp = ld2(P.phi + growp + x), zm = ld2(P.phi + s - plane), zp = ld2(P.phi + s + plane);
					const double pm = P.phi[grow + xm], pp = P.phi[grow + xp];
					const double2 src = ld2s(P.src + s), ch = ld2s(P.chi + s);
					acc += src.x; acc += src.y;
					z[m].x = scalar_site(src.x, p.x, ch.x, pm, p.y, ym.x, yp.x, zm.x, zp.x, P.bgmodel, P.coeff, P.coeff2, P.coeff3);
					z[m].y = scalar_site(src.y, p.y, ch.y, p.x, pp, ym.y, yp.y, zm.y, zp.y, P.bgmodel, P.coeff, P.coeff2, P.coeff3);
				}
				else
				{
					// gevolution.hpp:64-143, the operations of tensor_site() for this warp's component only
					const double2 t = ld2s(P.src + comp * P.cs_src + s);
					double va = P.coeff * t.x, vb = P.coeff * t.y;
					if (comp == 0)
					{
						const double2 p = ld2(P.phi + s);
						const double pm = P.phi[grow + xm], pp = P.phi[grow + xp];
						va += 0.5 * (p.y - pm) * (p.y - pm); vb += 0.5 * (pp - p.x) * (pp - p.x);
					}
					else if (comp == 3)
					{
						const double2 yp = ld2(P.phi + growp + x), ym = ld2(P.phi + growm + x);
						va += 0.5 * (yp.x - ym.x) * (yp.x - ym.x); vb += 0.5 * (yp.y - ym.y) * (yp.y - ym.y);
					}
					else if (comp == 5)
					{
						const double2 zp = ld2(P.phi + s + plane), zm = ld2(P.phi + s - plane);
						va += 0.5 * (zp.x - zm.x) * (zp.x - zm.x); vb += 0.5 * (zp.y - zm.y) * (zp.y - zm.y);
					}
					else
					{
						// mixed components (i, j): p0, the forward neighbours along i and j and the diagonal one
						const double2 p = ld2(P.phi + s);
						double ai, aj, aij, bi, bj, bij;                                // site a = x, site b = x + 1
						if (comp == 1)
						{
							const double2 yp = ld2(P.phi + growp + x);
							const double pp = P.phi[grow + xp], pyp = P.phi[growp + xp];
							ai = p.y; aj = yp.x; aij = yp.y; bi = pp; bj = yp.y; bij = pyp;
						}
						else if (comp == 2)
						{
							const double2 zp = ld2(P.phi + s + plane);
							const double pp = P.phi[grow + xp], pzp = P.phi[grow + xp + plane];
							ai = p.y; aj = zp.x; aij = zp.y; bi = pp; bj = zp.y; bij = pzp;
						}
						else
						{
							const double2 yp = ld2(P.phi + growp + x), zp = ld2(P.phi + s + plane), yz = ld2(P.phi + growp + x + plane);
							ai = yp.x; aj = zp.x; aij = yz.x; bi = yp.y; bj = zp.y; bij = yz.y;
						}
						va += ai * aj - p.x * aij;
						va += 0.5 * p.x * p.x; va -= 0.5 * ai * ai; va -= 0.5 * aj * aj; va += 0.5 * aij * aij;
						vb += bi * bj - p.y * bij;
						vb += 0.5 * p.y * p.y; vb -= 0.5 * bi * bi; vb -= 0.5 * bj * bj; vb += 0.5 * bij * bij;
					}
					z[m] = make_double2(va, vb);
				}
			}
			if (MODE == 1 && P.partial != NULL)
			{
				for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
				if (lane == 0) P.partial[row] = acc;
			}
		}
		row_fft_store(z, ws, tab, P.out + comp * P.cs_out + (size_t) row * XP_NH, lane);
	}
}

// y-pass of the forward transform, in place on [z][y][kx]: 512-point complex FFTs along y for eight adjacent kx columns per block.
// The tile (512 rows x 128 bytes) is brought in with the even and the odd rows of a column split (so that a warp reads its two
// 256-point halves contiguously), every warp transforms one column as two 256-point FFTs and the radix-2 step
// X[k] = E[k] + w512^k O[k], X[k + 256] = E[k] - w512^k O[k] (the un-packing table again), and the tile goes back row by row.
#define YP_COLS 8
#define YP_PITCH 521                                   // double2 per column in shared memory: 521 = 1 mod 8, so the eight columns of a row land in eight different 16-byte bank groups
__global__ void __launch_bounds__(YP_COLS * 32, 2) k_ypass(double2 * f, int nplanes, const double2 * __restrict__ gtab)
{
	extern __shared__ __align__(16) double2 xp_smem[];
	double2 * tab = xp_smem;
	double2 * cols = xp_smem + XP_TAB;
	for (int i = threadIdx.x; i < XP_TAB; i += YP_COLS * 32) tab[i] = gtab[i];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int kxl = threadIdx.x & (YP_COLS - 1), ys = threadIdx.x >> 3;          // tile traffic: 8 columns x 32 rows per step
	const int groups = (XP_NH + YP_COLS - 1) / YP_COLS;                             // 33: the last group holds the single column kx = 256
	const double2 * T3 = tab + 256 + 36;
	for (int t = blockIdx.x; t < nplanes * groups; t += gridDim.x)
	{
		const int zp = t / groups, kx0 = (t - zp * groups) * YP_COLS;
		const bool have = kx0 + kxl < XP_NH;
		double2 * plane = f + (size_t) zp * XP_N * XP_NH;
		__syncthreads();                                                            // the previous tile has been written out (and the tables are in)
		#pragma unroll 4
		for (int it = 0; it < XP_N / 32; it++)
		{
			const int y = it * 32 + ys;
			if (have) cols[kxl * YP_PITCH + (y & 1) * 256 + (y >> 1)] = __ldcs(plane + (size_t) y * XP_NH + kx0 + kxl);
		}
		__syncthreads();
		if (kx0 + warp < XP_NH)
		{
			double2 * col = cols + warp * YP_PITCH;
			double2 e[8], o[8];
			#pragma unroll
			for (int m = 0; m < 8; m++) { e[m] = col[lane + 32 * m]; o[m] = col[256 + lane + 32 * m]; }
			__syncwarp();
			fft256(e, col, tab, lane);
			__syncwarp();
			fft256(o, col, tab, lane);
			__syncwarp();
			#pragma unroll
			for (int m = 0; m < 8; m++)
			{
				const int k = lane + 32 * m;
				const double2 wo = cmul(T3[k], o[m]);
				col[k] = cadd(e[m], wo); col[k + 256] = csub(e[m], wo);
			}
		}
		__syncthreads();
		#pragma unroll 4
		for (int it = 0; it < XP_N / 32; it++)
		{
			const int y = it * 32 + ys;
			if (have) plane[(size_t) y * XP_NH + kx0 + kxl] = cols[kxl * YP_PITCH + y];
		}
	}
}

__global__ void k_sum_rows(const double * __restrict__ partial, int n, double * out)
{
	__shared__ double sh[1024];
	double t = 0.;
	for (int i = threadIdx.x; i < n; i += 1024) t += partial[i];
	sh[threadIdx.x] = t;
	__syncthreads();
	for (int o = 512; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) *out = sh[0];
}

// twiddle tables, once per context (kept in the context's k-table allocation list through a static per-device cache)
double2 * g_tab[64] = {NULL};
int table_for(gevb_ctx * c, const double2 ** out)
{
	GEVB_CHECK_ARG(c->device >= 0 && c->device < 64, "xpass: device index out of range");
	if (g_tab[c->device] == NULL)
	{
		std::vector<double2> h(XP_TAB);
		auto w = [](int N, int e) { double s, co; sincospi(-2.0 * (double) (e % N) / (double) N, &s, &co); return make_double2(co, s); };
		for (int cc = 0; cc < 8; cc++) for (int b = 0; b < 32; b++) h[cc * 32 + b] = w(256, b * cc);
		for (int f = 0; f < 4; f++) for (int g = 0; g < 9; g++) h[256 + f * 9 + g] = w(32, f * (g < 8 ? g : 0));
		for (int k = 0; k < 256; k++) h[256 + 36 + k] = w(512, k);
		CUDA_TRY(cudaMalloc(&g_tab[c->device], XP_TAB * sizeof(double2)));
		CUDA_TRY(cudaMemcpy(g_tab[c->device], h.data(), XP_TAB * sizeof(double2), cudaMemcpyHostToDevice));
	}
	*out = g_tab[c->device];
	return 0;
}

template <int MODE>
int launch_xpass(gevb_ctx * c, const XParams & P)
{
	constexpr int WARPS = MODE == 2 ? XP_WARPS_TENSOR : XP_WARPS_PLAIN;
	const size_t smem = (XP_TAB + (size_t) WARPS * XP_WS) * sizeof(double2);
	CUDA_TRY(cudaFuncSetAttribute(k_xpass<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int per_sm = 0;
	CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_xpass<MODE>, WARPS * 32, smem));
	const int rows_per_block = MODE == 2 ? WARPS / 6 : WARPS;
	const int want = (P.nzl * XP_N + rows_per_block - 1) / rows_per_block, persistent = c->num_sms * (per_sm > 0 ? per_sm : 1);
	k_xpass<MODE><<<want < persistent ? want : persistent, WARPS * 32, smem, c->stream>>>(P);
	KERNEL_CHECK(c);
	return 0;
}

} // namespace

// is the own x-pass usable for this plan (single rank, N = 512, knob on)?
bool gevb_xpass_available(const gevb_plan * p)
{
	return p != NULL && !p->multi && p->ctx->N == XP_N && gevb_tune(TUNE_FFT_XPASS) != 0 && gevb_tune(TUNE_FFT_DECOMPOSED) != 0;
}

// forward transform of the plan's field through the own x-pass.  mode 0: the real field as it is; 1: scalar prepareFTsource fused
// (the plan's real field is the incoming source and is left as it is; sum_dev != NULL receives the sum of the incoming source);
// 2: tensor prepareFTsource fused (T = the plan's real field, left as it is).
int gevb_xpass_forward(gevb_plan * p, int mode, const double * phi, const double * chi, double bgmodel, double coeff, double coeff2, double coeff3, double * sum_dev)
{
	gevb_ctx * c = p->ctx;
	gevb_field * rf = p->real_field, * cf = p->cplx_field;
	const int N = c->N, nh = c->nh, nc = rf->ncomp;
	XParams P;
	memset(&P, 0, sizeof(P));
	P.nzl = c->nzl; P.phi = phi; P.chi = chi; P.bgmodel = bgmodel; P.coeff = coeff; P.coeff2 = coeff2; P.coeff3 = coeff3;
	P.out = (double2 *) cf->data; P.cs_out = cf->comp_stride; P.cs_src = rf->comp_stride;
	GEVB_TRY(table_for(c, &P.tab));
	if (mode == 2)
	{
		GEVB_CHECK_ARG(nc == 6, "xpass: the tensor source has six components");
		P.src = rf->data;                                         // ghosted addressing
		GEVB_TRY(launch_xpass<2>(c, P));
	}
	else if (mode == 1)
	{
		void * part = NULL;
		if (sum_dev) GEVB_TRY(gevb_ctx_scratch(c, (size_t) c->nzl * N * sizeof(double), &part));
		P.src = rf->data; P.partial = (double *) part;
		GEVB_TRY(launch_xpass<1>(c, P));
		if (sum_dev) { k_sum_rows<<<1, 1024, 0, c->stream>>>((const double *) part, c->nzl * N, sum_dev); KERNEL_CHECK(c); }
	}
	else
		for (int k = 0; k < nc; k++)
		{
			P.src = rf->data + c->plane() + k * rf->comp_stride;      // bulk: the row index counts from the first owned plane
			P.out = (double2 *) cf->data + k * cf->comp_stride;
			GEVB_TRY(launch_xpass<0>(c, P));
		}
	{
		// y-pass: own kernel, in place; z-pass: cuFFT's 1-D transform along z (the plan the cuFFT path uses too)
		const size_t smem = (XP_TAB + (size_t) YP_COLS * YP_PITCH) * sizeof(double2);
		CUDA_TRY(cudaFuncSetAttribute(k_ypass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		const int tiles = c->nzl * ((XP_NH + YP_COLS - 1) / YP_COLS), persistent = c->num_sms * 2;
		for (int k = 0; k < nc; k++)
		{
			double2 * f = (double2 *) cf->data + k * cf->comp_stride;
			k_ypass<<<tiles < persistent ? tiles : persistent, YP_COLS * 32, smem, c->stream>>>(f, c->nzl, P.tab);
			KERNEL_CHECK(c);
			CUFFT_TRY(cufftExecZ2Z(p->bz1d, (cufftDoubleComplex *) f, (cufftDoubleComplex *) f, CUFFT_FORWARD));
		}
	}
	c->launches += 3 * nc;
	return 0;
}
