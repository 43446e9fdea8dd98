// spectrum.cu -- binned power spectrum of a Fourier-space field
//
//   extractCrossSpectrum with fld1 == fld2 / extractPowerSpectrum   tools.hpp:53-240
//
// One pass over the local k-sites; per-warp aggregation of equal bins
// (__match_any_sync) followed by FP64 reductions into the five bin arrays,
// NCCL all-reduce across ranks (MPI_Reduce in the reference, tools.hpp:169-211),
// final normalisation on the host exactly as tools.hpp:186-193.
#include <math.h>
#include "gevb_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_spectrum(KLayout L, const double * __restrict__ typek2, const double * __restrict__ sinc,
	const double2 * __restrict__ F, size_t cs, int ncomp, int symm, int numbins, double k2max,
	double * kbin, double * power, double * kscatter, double * pscatter, unsigned long long * occupation)
{
	const size_t stride = (size_t) gridDim.x * blockDim.x;
	const size_t rounds = (L.sites + stride - 1) / stride;
	for (size_t it = 0; it < rounds; it++)
	{
		const size_t i = it * stride + blockIdx.x * (size_t) blockDim.x + threadIdx.x;
		int bin = -1, weight = 0;
		double k2 = 0., p = 0., s = 1.;
		if (i < L.sites)
		{
			int kx, ky, kz; k_decode(L, i, kx, ky, kz);
			if ((kx | ky | kz) != 0)                                                    // tools.hpp:121-122
			{
				weight = (kx == 0 || (kx == L.N / 2 && L.N % 2 == 0)) ? 1 : 2;          // :123-128
				k2 = typek2[kx] + typek2[ky] + typek2[kz];                              // :130
				s = sinc[kx] * sinc[ky] * sinc[kz]; s *= s;                             // :131-132
				if (symm)
				{                                                                       // :138-147
					const double2 a1 = F[cs + i], a2 = F[2 * cs + i], a4 = F[4 * cs + i], a0 = F[i], a3 = F[3 * cs + i], a5 = F[5 * cs + i];
					p = a1.x * a1.x + a1.y * a1.y; p += a2.x * a2.x + a2.y * a2.y; p += a4.x * a4.x + a4.y * a4.y;
					p *= 2.;
					p += a0.x * a0.x + a0.y * a0.y; p += a3.x * a3.x + a3.y * a3.y; p += a5.x * a5.x + a5.y * a5.y;
				}
				else
					for (int c = 0; c < ncomp; c++) { const double2 v = F[c * cs + i]; p += v.x * v.x + v.y * v.y; }   // :149-153
				bin = (int) floor((double) numbins * sqrt(k2 / k2max));                // :155
				if (bin >= numbins) bin = -1;
			}
		}
		double v0 = weight * sqrt(k2), v1 = weight * k2, v2 = weight * p * k2 * sqrt(k2) / s, v3 = weight * p * p * k2 * k2 * k2 / s / s;   // :158-161
		unsigned long long v4 = weight;
		// lanes with the same bin: sum within the group, leader issues the reductions
		const unsigned group = __match_any_sync(0xffffffffu, bin);
		const int leader = __ffs(group) - 1, lane = threadIdx.x & 31;
		for (unsigned m = group & ~(1u << leader); m; m &= m - 1)
		{
			const int src = __ffs(m) - 1;
			const double t0 = __shfl_sync(group, v0, src), t1 = __shfl_sync(group, v1, src), t2 = __shfl_sync(group, v2, src), t3 = __shfl_sync(group, v3, src);
			const unsigned long long t4 = __shfl_sync(group, v4, src);
			if (lane == leader) { v0 += t0; v1 += t1; v2 += t2; v3 += t3; v4 += t4; }
		}
		if (lane == leader && bin >= 0)
		{
			atomicAdd(kbin + bin, v0); atomicAdd(kscatter + bin, v1); atomicAdd(power + bin, v2); atomicAdd(pscatter + bin, v3);
			atomicAdd(occupation + bin, v4);
		}
	}
}

} // namespace

extern "C" int gevb_extractPowerSpectrum(gevb_field * f, double * kbin, double * power, double * kscatter, double * pscatter, int * occupation, int numbins, int deconvolve, int ktype)
{
	GEVB_CHECK_ARG(f != NULL && f->kind == GEVB_CPLX, "extractPowerSpectrum: needs a Fourier-space field");
	GEVB_CHECK_ARG(kbin && power && kscatter && pscatter && occupation, "extractPowerSpectrum: NULL output array");
	GEVB_CHECK_ARG(numbins >= 1 && numbins <= 1 << 20, "extractPowerSpectrum: bad number of bins %d", numbins);
	gevb_ctx * c = f->ctx;
	CUDA_TRY(cudaSetDevice(c->device));
	Timed timed_(c, CLS_SPECTRUM);
	const int N = c->N;
	// tables exactly as tools.hpp:64-106
	std::vector<double> tab(2 * N);
	double * typek2 = tab.data(), * sinc = tab.data() + N;
	int i;
	if (ktype == 0) { for (i = 0; i < N; i++) { typek2[i] = 2. * (double) N * sin(M_PI * (double) i / (double) N); typek2[i] *= typek2[i]; } }
	else
	{
		for (i = 0; i <= N / 2; i++) { typek2[i] = 2. * M_PI * (double) i; typek2[i] *= typek2[i]; }
		for (; i < N; i++) { typek2[i] = 2. * M_PI * (double) (N - i); typek2[i] *= typek2[i]; }
	}
	sinc[0] = 1.;
	for (i = 1; i <= N / 2; i++) sinc[i] = deconvolve ? sin(M_PI * (float) i / (float) N) * (float) N / (M_PI * (float) i) : 1.;
	for (; i < N; i++) sinc[i] = sinc[N - i];
	const double k2max = 3. * typek2[N / 2];                                            // :108
	void * buf;
	const size_t nb = (size_t) numbins;
	GEVB_TRY(gevb_ctx_scratch(c, (2 * N + 5 * nb) * sizeof(double), &buf));
	double * d_tab = (double *) buf, * d_bins = d_tab + 2 * N;
	CUDA_TRY(cudaMemcpyAsync(d_tab, tab.data(), 2 * N * sizeof(double), cudaMemcpyHostToDevice, c->stream));
	CUDA_TRY(cudaMemsetAsync(d_bins, 0, 5 * nb * sizeof(double), c->stream));
	KLayout L = make_klayout(c);
	k_spectrum<<<gevb_grid(c, L.sites, 256), 256, 0, c->stream>>>(L, d_tab, d_tab + N, (const double2 *) f->data, f->comp_stride, f->ncomp, f->symmetric, numbins, k2max,
		d_bins, d_bins + nb, d_bins + 2 * nb, d_bins + 3 * nb, (unsigned long long *) (d_bins + 4 * nb));
	KERNEL_CHECK(c);
	if (c->nranks > 1)
	{
		NCCL_TRY(ncclAllReduce(d_bins, d_bins, 4 * nb, ncclDouble, ncclSum, c->comm, c->stream));
		NCCL_TRY(ncclAllReduce(d_bins + 4 * nb, d_bins + 4 * nb, nb, ncclUint64, ncclSum, c->comm, c->stream));
	}
	std::vector<double> h(5 * nb);
	CUDA_TRY(cudaMemcpyAsync(h.data(), d_bins, 5 * nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	const unsigned long long * occ = (const unsigned long long *) (h.data() + 4 * nb);
	for (i = 0; i < numbins; i++)
	{
		kbin[i] = h[i]; power[i] = h[nb + i]; kscatter[i] = h[2 * nb + i]; pscatter[i] = h[3 * nb + i]; occupation[i] = (int) occ[i];
		if (occupation[i] > 0)
		{                                                                               // :186-193
			kscatter[i] = sqrt(kscatter[i] * occupation[i] - kbin[i] * kbin[i]) / occupation[i];
			if (!isfinite(kscatter[i])) kscatter[i] = 0.;
			kbin[i] = kbin[i] / occupation[i];
			power[i] /= occupation[i];
			pscatter[i] = sqrt(pscatter[i] / occupation[i] - power[i] * power[i]);
			if (!isfinite(pscatter[i])) pscatter[i] = 0.;
		}
	}
	return 0;
}
