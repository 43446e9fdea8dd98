// source_kernels.cu -- real-space source preparation, one HBM pass each
//
//   prepareFTsource (scalar)  gevolution.hpp:170-192   32 B / site (phi, chi, source in; result out)
//   prepareFTsource (tensor)  gevolution.hpp:57-147   104 B / site (phi, 6 Tij in; 6 Sij out)
//
// PHINONLINEAR on, ORIGINALMETRIC off (reference makefile:21).  One thread per
// pair of x-adjacent sites, x fastest so every load is a coalesced row segment; the phi stencil
// neighbours (x+-1 in the row, y+-1 rows, z+-1 planes incl. ghost planes) are
// served from L1/L2, so DRAM sees each phi value once.  x and y wrap by index
// arithmetic; z uses the ghost planes filled by updateHalo.
#include "gevb_internal.cuh"

namespace {

struct RGeom { int N, nzl; size_t plane; uint64_t mhalf, mN; };   // mhalf = ceil(2^40 / (N/2)), mN = ceil(2^40 / N): quotients by multiply-shift
__host__ __device__ __forceinline__ uint32_t mdiv(uint32_t i, uint64_t m) { return (uint32_t) (((uint64_t) i * m) >> 40); }
static RGeom make_rgeom(const gevb_ctx * c)
{
	RGeom G = {c->N, c->nzl, c->plane(), ((1ull << 40) + (uint64_t) (c->N / 2) - 1) / (uint64_t) (c->N / 2), ((1ull << 40) + (uint64_t) c->N - 1) / (uint64_t) c->N};
	return G;
}

// Two x-adjacent sites per thread: every stream (phi rows, the six T / S components, source, chi) moves as 16-byte
// accesses, which halves the number of memory requests in flight per byte; the arithmetic per site is unchanged.
// N is even (checked at context creation), so a pair never straddles a row and every double2 is aligned.
__device__ __forceinline__ double2 ld2(const double * p) { return *(const double2 *) p; }
__device__ __forceinline__ double2 ld2s(const double * p) { return __ldcs((const double2 *) p); }
__device__ __forceinline__ void st2s(double * p, double a, double b) { __stcs((double2 *) p, make_double2(a, b)); }

// gevolution.hpp:176-190 for one site
__device__ __forceinline__ double scalar_site(double src, double p, double chi, double pxm, double pxp, double pym, double pyp, double pzm, double pzp,
                                              double bgmodel, double coeff, double coeff2, double coeff3)
{
	double res = coeff2 * (src - bgmodel);                                          // :176
	res *= 1. - 2. * p;                                                             // :184
	const double d0 = pxm - pxp, d1 = pym - pyp, d2 = pzm - pzp;
	res += 0.125 * d0 * d0;                                                         // :185
	res += 0.125 * d1 * d1;                                                         // :186
	res += 0.125 * d2 * d2;                                                         // :187
	res += (coeff3 - coeff) * p - coeff3 * chi;                                     // :190
	return res;
}

// partial: per-block sums of the incoming source (main.cpp:459-462 fused into this pass), or NULL
__global__ void __launch_bounds__(256) k_prepare_scalar(RGeom G, const double * __restrict__ phi, const double * __restrict__ chi, const double * source, double bgmodel, double * result,
                                                        double coeff, double coeff2, double coeff3, double * partial)
{
	const int N = G.N, half = N >> 1;
	const size_t pairs = (size_t) G.nzl * G.plane / 2;
	double acc = 0.;
	for (size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x; j < pairs; j += (size_t) gridDim.x * blockDim.x)
	{
		const uint32_t r = mdiv((uint32_t) j, G.mhalf), zq = mdiv(r, G.mN);          // pairs < 2^31 (checked by the launcher)
		const int x = 2 * (int) ((uint32_t) j - r * (uint32_t) half);
		const int y = (int) (r - zq * (uint32_t) N); const int zl = (int) zq;
		const size_t row = ((size_t) (zl + 1) * N + y) * N;
		const int xm = x == 0 ? N - 1 : x - 1, xp = x == N - 2 ? 0 : x + 2;
		const size_t rowm = ((size_t) (zl + 1) * N + (y == 0 ? N - 1 : y - 1)) * N;
		const size_t rowp = ((size_t) (zl + 1) * N + (y == N - 1 ? 0 : y + 1)) * N;
		const size_t s = row + x;
		const double2 p = ld2(phi + s), ym = ld2(phi + rowm + x), yp = ld2(phi + rowp + x), zm = ld2(phi + s - G.plane), zp = ld2(phi + s + G.plane);
		const double pm = phi[row + xm], pp = phi[row + xp];
		const double2 src = ld2s(source + s), ch = ld2s(chi + s);
		acc += src.x; acc += src.y;
		const double ra = scalar_site(src.x, p.x, ch.x, pm, p.y, ym.x, yp.x, zm.x, zp.x, bgmodel, coeff, coeff2, coeff3);
		const double rb = scalar_site(src.y, p.y, ch.y, p.x, pp, ym.y, yp.y, zm.y, zp.y, bgmodel, coeff, coeff2, coeff3);
		st2s(result + s, ra, rb);
	}
	if (partial != NULL)
	{
		__shared__ double wsum[8];
		for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
		if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			double t = 0.;
			for (int w = 0; w < 8; w++) t += wsum[w];
			partial[blockIdx.x] = t;
		}
	}
}

__global__ void k_sum_partials(const double * __restrict__ partial, int n, double * out)
{
	__shared__ double sh[256];
	double t = 0.;
	for (int i = threadIdx.x; i < n; i += 256) t += partial[i];
	sh[threadIdx.x] = t;
	__syncthreads();
	for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
	if (threadIdx.x == 0) *out = sh[0];
}

// gevolution.hpp:64-143 for one site: the six S components in field order (0,0),(0,1),(0,2),(1,1),(1,2),(2,2)
struct PhiStencil { double p0, px, mx, py, my, pz, mz, pxy, pxz, pyz; };
__device__ __forceinline__ void tensor_site(const PhiStencil & f, const double * t, double coeff, double * o)
{
	double v;
	v = coeff * t[0]; v += 0.5 * (f.px - f.mx) * (f.px - f.mx); o[0] = v;                                      // :64,70
	v = coeff * t[3]; v += 0.5 * (f.py - f.my) * (f.py - f.my); o[3] = v;                                      // :75,81
	v = coeff * t[5]; v += 0.5 * (f.pz - f.mz) * (f.pz - f.mz); o[5] = v;                                      // :86,92
	v = coeff * t[1];                                                                                         // :97
	v += f.px * f.py - f.p0 * f.pxy;                                                                          // :99
	v += 0.5 * f.p0 * f.p0; v -= 0.5 * f.px * f.px; v -= 0.5 * f.py * f.py; v += 0.5 * f.pxy * f.pxy;          // :106-109
	o[1] = v;
	v = coeff * t[2];                                                                                         // :114
	v += f.px * f.pz - f.p0 * f.pxz;                                                                          // :116
	v += 0.5 * f.p0 * f.p0; v -= 0.5 * f.px * f.px; v -= 0.5 * f.pz * f.pz; v += 0.5 * f.pxz * f.pxz;          // :123-126
	o[2] = v;
	v = coeff * t[4];                                                                                         // :131
	v += f.py * f.pz - f.p0 * f.pyz;                                                                          // :133
	v += 0.5 * f.p0 * f.p0; v -= 0.5 * f.py * f.py; v -= 0.5 * f.pz * f.pz; v += 0.5 * f.pyz * f.pyz;          // :140-143
	o[4] = v;
}

__global__ void __launch_bounds__(256) k_prepare_tensor(RGeom G, const double * __restrict__ phi, const double * T, double * S, size_t cs, double coeff)
{
	const int N = G.N, half = N >> 1;
	const size_t pairs = (size_t) G.nzl * G.plane / 2;
	const size_t pl = G.plane;
	for (size_t j = blockIdx.x * (size_t) blockDim.x + threadIdx.x; j < pairs; j += (size_t) gridDim.x * blockDim.x)
	{
		const uint32_t r = mdiv((uint32_t) j, G.mhalf), zq = mdiv(r, G.mN);          // pairs < 2^31 (checked by the launcher)
		const int x = 2 * (int) ((uint32_t) j - r * (uint32_t) half);
		const int y = (int) (r - zq * (uint32_t) N); const int zl = (int) zq;
		const int xm = x == 0 ? N - 1 : x - 1, xp = x == N - 2 ? 0 : x + 2;
		const int ym = y == 0 ? N - 1 : y - 1, yp = y == N - 1 ? 0 : y + 1;
		const size_t row = ((size_t) (zl + 1) * N + y) * N, rowm = ((size_t) (zl + 1) * N + ym) * N, rowp = ((size_t) (zl + 1) * N + yp) * N;
		const size_t s = row + x;
		double2 t[6];
		#pragma unroll
		for (int c = 0; c < 6; c++) t[c] = ld2s(T + c * cs + s);
		const double2 p = ld2(phi + s), vyp = ld2(phi + rowp + x), vym = ld2(phi + rowm + x), vzp = ld2(phi + s + pl), vzm = ld2(phi + s - pl), vyz = ld2(phi + rowp + x + pl);
		const double pm = phi[row + xm], pp = phi[row + xp], pyp = phi[rowp + xp], pzp = phi[row + xp + pl];
		const PhiStencil fa = {p.x, p.y, pm, vyp.x, vym.x, vzp.x, vzm.x, vyp.y, vzp.y, vyz.x};
		const PhiStencil fb = {p.y, pp, p.x, vyp.y, vym.y, vzp.y, vzm.y, pyp, pzp, vyz.y};
		double ta[6], tb[6], oa[6], ob[6];
		#pragma unroll
		for (int c = 0; c < 6; c++) { ta[c] = t[c].x; tb[c] = t[c].y; }
		tensor_site(fa, ta, coeff, oa);
		tensor_site(fb, tb, coeff, ob);
		#pragma unroll
		for (int c = 0; c < 6; c++) st2s(S + c * cs + s, oa[c], ob[c]);
	}
}

int check_real(const gevb_field * f, int ncomp, const char * who, const char * name)
{
	GEVB_CHECK_ARG(f != NULL, "%s: %s is NULL", who, name);
	GEVB_CHECK_ARG(f->kind == GEVB_REAL, "%s: %s must be a real-space field", who, name);
	GEVB_CHECK_ARG(f->ncomp == ncomp, "%s: %s needs %d components (has %d)", who, name, ncomp, f->ncomp);
	return 0;
}

} // namespace

static int prepare_scalar(gevb_field * phi, gevb_field * chi, gevb_field * source, double bgmodel, gevb_field * result, double coeff, double coeff2, double coeff3, double * sum_source)
{
	GEVB_TRY(check_real(phi, 1, "prepareFTsource", "phi"));
	GEVB_TRY(check_real(chi, 1, "prepareFTsource", "chi"));
	GEVB_TRY(check_real(source, 1, "prepareFTsource", "source"));
	GEVB_TRY(check_real(result, 1, "prepareFTsource", "result"));
	GEVB_CHECK_ARG(result != phi && result != chi, "prepareFTsource: result must not alias phi or chi");
	gevb_ctx * c = phi->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_PREP_SCALAR);
	RGeom G = make_rgeom(c);
	GEVB_CHECK_ARG((size_t) c->nzl * c->plane() / 2 < (1ull << 31), "prepareFTsource: too many sites per rank for the 32-bit index decode (use more ranks)");
	const int grid = gevb_grid(c, (size_t) c->nzl * c->plane() / 2, 256);      // <= 8 blocks per SM: the partials fit d_red[0..2048)
	GEVB_CHECK_ARG(sum_source == NULL || grid <= 2048, "prepareFTsource: reduction buffer too small for %d blocks", grid);
	k_prepare_scalar<<<grid, 256, 0, c->stream>>>(G, phi->data, chi->data, source->data, bgmodel, result->data, coeff, coeff2, coeff3, sum_source ? c->d_red : NULL);
	KERNEL_CHECK(c);
	if (sum_source)
	{
		k_sum_partials<<<1, 256, 0, c->stream>>>(c->d_red, grid, c->d_red + 2048);
		KERNEL_CHECK(c);
		if (c->nranks > 1) NCCL_TRY(ncclAllReduce(c->d_red + 2048, c->d_red + 2048, 1, ncclDouble, ncclSum, c->comm, c->stream));
		CUDA_TRY(cudaMemcpyAsync(c->h_red, c->d_red + 2048, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*sum_source = c->h_red[0];
	}
	return 0;
}

extern "C" int gevb_prepareFTsource_scalar(gevb_field * phi, gevb_field * chi, gevb_field * source, double bgmodel, gevb_field * result, double coeff, double coeff2, double coeff3)
{
	return prepare_scalar(phi, chi, source, bgmodel, result, coeff, coeff2, coeff3, NULL);
}

extern "C" int gevb_prepareFTsource_scalar_sum(gevb_field * phi, gevb_field * chi, gevb_field * source, double bgmodel, gevb_field * result, double coeff, double coeff2, double coeff3, double * sum_source)
{
	GEVB_CHECK_ARG(sum_source != NULL, "prepareFTsource: sum_source is NULL");
	return prepare_scalar(phi, chi, source, bgmodel, result, coeff, coeff2, coeff3, sum_source);
}

extern "C" int gevb_prepareFTsource_tensor(gevb_field * phi, gevb_field * Tij, gevb_field * Sij, double coeff)
{
	GEVB_TRY(check_real(phi, 1, "prepareFTsource", "phi"));
	GEVB_TRY(check_real(Tij, 6, "prepareFTsource", "Tij"));
	GEVB_TRY(check_real(Sij, 6, "prepareFTsource", "Sij"));
	gevb_ctx * c = phi->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_PREP_TENSOR);
	RGeom G = make_rgeom(c);
	GEVB_CHECK_ARG((size_t) c->nzl * c->plane() / 2 < (1ull << 31), "prepareFTsource: too many sites per rank for the 32-bit index decode (use more ranks)");
	k_prepare_tensor<<<gevb_grid(c, (size_t) c->nzl * c->plane() / 2, 256), 256, 0, c->stream>>>(G, phi->data, Tij->data, Sij->data, Sij->comp_stride, coeff);
	KERNEL_CHECK(c);
	return 0;
}

// fused with the forward transform (xpass.cu) where that applies
extern "C" int gevb_prepareFTsource_scalar_fft(gevb_field * phi, gevb_field * chi, gevb_plan * plan, double bgmodel, double coeff, double coeff2, double coeff3, double * sum_source)
{
	GEVB_CHECK_ARG(plan != NULL, "prepareFTsource: NULL plan");
	gevb_field * source = plan->real_field;
	if (!gevb_xpass_available(plan))
	{
		GEVB_TRY(prepare_scalar(phi, chi, source, bgmodel, source, coeff, coeff2, coeff3, sum_source));
		return gevb_plan_execute(plan, GEVB_FFT_FORWARD);
	}
	GEVB_TRY(check_real(phi, 1, "prepareFTsource", "phi"));
	GEVB_TRY(check_real(chi, 1, "prepareFTsource", "chi"));
	GEVB_TRY(check_real(source, 1, "prepareFTsource", "source"));
	GEVB_CHECK_ARG(source != phi && source != chi, "prepareFTsource: result must not alias phi or chi");
	gevb_ctx * c = phi->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_FFT_FWD);
	GEVB_TRY(gevb_xpass_forward(plan, 1, phi->data, chi->data, bgmodel, coeff, coeff2, coeff3, sum_source ? c->d_red + 2048 : NULL));
	if (sum_source)
	{
		CUDA_TRY(cudaMemcpyAsync(c->h_red, c->d_red + 2048, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		*sum_source = c->h_red[0];
	}
	return 0;
}

extern "C" int gevb_prepareFTsource_tensor_fft(gevb_field * phi, gevb_plan * plan, double coeff)
{
	GEVB_CHECK_ARG(plan != NULL, "prepareFTsource: NULL plan");
	gevb_field * Sij = plan->real_field;
	if (!gevb_xpass_available(plan))
	{
		GEVB_TRY(gevb_prepareFTsource_tensor(phi, Sij, Sij, coeff));
		return gevb_plan_execute(plan, GEVB_FFT_FORWARD);
	}
	GEVB_TRY(check_real(phi, 1, "prepareFTsource", "phi"));
	GEVB_TRY(check_real(Sij, 6, "prepareFTsource", "Tij"));
	gevb_ctx * c = phi->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_FFT_FWD);
	return gevb_xpass_forward(plan, 2, phi->data, NULL, 0., coeff, 0., 0., NULL);
}
