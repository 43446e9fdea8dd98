// timing.cu -- optional per-call device timing (CUDA events on the context's stream)
//
// bench.py switches this on for the timed region to obtain the live per-kernel
// durations that the roofline figures are computed from.  Each C-ABI call that
// enqueues work records an event pair around it; gevb_ctx_timing_read() drains
// the pairs into per-class totals.  Off by default (no events recorded).
#include "gevb_internal.cuh"

static const char * k_names[GEVB_NCLS] = {
	"projection_init", "projection_T00_project", "projection_Tij_project", "projection_T00_Tij_project", "projection_T0i_project",
	"projection_comm", "field_sum", "prepareFTsource_scalar", "prepareFTsource_tensor", "fft_forward", "fft_backward",
	"solveModifiedPoissonFT", "projectFTscalar", "evolveFTvector", "projectFTvector", "projectFTtensor", "updateHalo",
	"updateVel", "moveParticles", "kick_drift", "rebin_sort", "extractPowerSpectrum", "migrate", "fft_alltoall", "fft_transpose",
	"projectFTscalar_evolveFTvector"
};

extern "C" const char * gevb_timing_class_name(int cls) { return (cls >= 0 && cls < GEVB_NCLS) ? k_names[cls] : NULL; }
extern "C" int gevb_timing_num_classes(void) { return GEVB_NCLS; }

extern "C" int gevb_ctx_timing(gevb_ctx * c, int enable)
{
	GEVB_CHECK_ARG(c != NULL, "gevb_ctx_timing: NULL context");
	if (c->timer == NULL) c->timer = new GevbTimer();
	c->timer->on = enable != 0;
	return 0;
}

void gevb_timer_begin(gevb_ctx * c, int cls)
{
	GevbTimer * t = c->timer;
	if (t == NULL || !t->on) return;
	if (t->used + 2 > t->ev.size())
	{
		size_t old = t->ev.size();
		t->ev.resize(old + 256);
		for (size_t i = old; i < t->ev.size(); i++) cudaEventCreate(&t->ev[i]);
	}
	t->cls.push_back(cls);
	cudaEventRecord(t->ev[t->used], c->stream);
	t->open.push_back(t->used);
	t->used += 2;
}

void gevb_timer_end(gevb_ctx * c)
{
	GevbTimer * t = c->timer;
	if (t == NULL || !t->on || t->open.empty()) return;
	cudaEventRecord(t->ev[t->open.back() + 1], c->stream);
	t->open.pop_back();
}

// totals since the last read: milliseconds and call counts per class (arrays of gevb_timing_num_classes())
extern "C" int gevb_ctx_timing_read(gevb_ctx * c, double * ms, int64_t * counts)
{
	GEVB_CHECK_ARG(c != NULL && ms != NULL && counts != NULL, "gevb_ctx_timing_read: NULL argument");
	for (int i = 0; i < GEVB_NCLS; i++) { ms[i] = 0.; counts[i] = 0; }
	GevbTimer * t = c->timer;
	if (t == NULL) return 0;
	CUDA_TRY(cudaStreamSynchronize(c->stream));
	for (size_t k = 0; k < t->cls.size(); k++)
	{
		float e = 0.f;
		CUDA_TRY(cudaEventElapsedTime(&e, t->ev[2 * k], t->ev[2 * k + 1]));
		ms[t->cls[k]] += e; counts[t->cls[k]]++;
	}
	t->cls.clear(); t->open.clear(); t->used = 0;
	return 0;
}
