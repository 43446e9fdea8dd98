// nccl_dl.cu -- NCCL is bound at run time (dlopen), not at link time
//
// The host process usually has PyTorch loaded, which ships its own libnccl.so.2;
// linking a second copy by DT_NEEDED makes whichever loads first win and breaks
// the other.  So libgevb.so resolves the few NCCL entry points it needs lazily,
// the first time a multi-rank context is created: the copy already mapped into
// the process is preferred (RTLD_NOLOAD), then $GEVB_NCCL_LIB, then the default
// search path.  Single-rank use never touches NCCL.
#include <dlfcn.h>
#include <stdlib.h>
#include "gevb_internal.cuh"

static GevbNccl g_nccl;
static bool g_loaded = false;

GevbNccl * gevb_nccl() { return &g_nccl; }

int gevb_nccl_load()
{
	if (g_loaded) return 0;
	void * h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
	if (h == NULL)
	{
		const char * env = getenv("GEVB_NCCL_LIB");
		if (env != NULL && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
	}
	if (h == NULL) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (h == NULL) GEVB_FAIL("NCCL: cannot load libnccl.so.2 (%s); set GEVB_NCCL_LIB", dlerror());
#define BIND(field, sym) do { *(void **) (&g_nccl.field) = dlsym(h, sym); if (g_nccl.field == NULL) GEVB_FAIL("NCCL: symbol %s not found", sym); } while (0)
	BIND(GetUniqueId, "ncclGetUniqueId");
	BIND(CommInitRank, "ncclCommInitRank");
	BIND(CommDestroy, "ncclCommDestroy");
	BIND(Send, "ncclSend");
	BIND(Recv, "ncclRecv");
	BIND(AllReduce, "ncclAllReduce");
	BIND(GroupStart, "ncclGroupStart");
	BIND(GroupEnd, "ncclGroupEnd");
	BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
	g_loaded = true;
	return 0;
}
