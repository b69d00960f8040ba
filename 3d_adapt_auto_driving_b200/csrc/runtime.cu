// runtime.cu -- library-wide state of the C-ABI: last error string, version.
#include "common.cuh"
#include <string.h>

static thread_local char g_last_error[256] = "";

void pn2_set_last_error(const char *msg) {
    strncpy(g_last_error, msg ? msg : "", sizeof(g_last_error) - 1);
    g_last_error[sizeof(g_last_error) - 1] = 0;
}

PN2_API const char *pn2_last_error(void) { return g_last_error; }
PN2_API int pn2_abi_version(void) { return 1; }
