// Error plumbing and small host utilities shared by all b200pci translation units.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace b200pci {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return B200PCI_ECUDA;
}

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;  // B200
    }
    return n;
}

}  // namespace b200pci

extern "C" int b200pci_version(void) { return B200PCI_VERSION; }
extern "C" const char *b200pci_last_error(void) { return b200pci::g_err; }
