// Error plumbing and ABI self-description of libgk_typing.so.
#include <stdarg.h>
#include <string.h>

#include "gk_common.cuh"

static thread_local char g_error[512] = "";

void gk_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

extern "C" const char* gk_last_error(void) { return g_error; }

extern "C" int gk_abi_version(void) { return 2; }

extern "C" int gk_sizeof(const char* name) {
    if (!strcmp(name, "GkMatrix")) return (int)sizeof(GkMatrix);
    if (!strcmp(name, "GkSearch")) return (int)sizeof(GkSearch);
    if (!strcmp(name, "GkLikItem")) return (int)sizeof(GkLikItem);
    if (!strcmp(name, "GkExpandItem")) return (int)sizeof(GkExpandItem);
    if (!strcmp(name, "GkScoreItem")) return (int)sizeof(GkScoreItem);
    if (!strcmp(name, "GkCountItem")) return (int)sizeof(GkCountItem);
    if (!strcmp(name, "GkPItem")) return (int)sizeof(GkPItem);
    if (!strcmp(name, "GkStepInfo")) return (int)sizeof(GkStepInfo);
    if (!strcmp(name, "GkEmProblem")) return (int)sizeof(GkEmProblem);
    return -1;
}
