/* Compiled as plain C11 by tests/test_abi.py: proves include/wax_vs_cuda.h is a valid C header and that the
   library links from C.  Exercises only the argument checks that need no device. */
#include <stdio.h>
#include <string.h>
#include "wax_vs_cuda.h"

int main(void) {
    wax_vs_engine *e = NULL;
    int32_t n = -1;
    if (wax_vs_create(0, WAX_VS_COSINE, NULL, 0, &e) != WAX_VS_ERR_ARGUMENT) return 1;
    if (strstr(wax_vs_last_error(), "dimensions must be > 0") == NULL) return 2;
    if (wax_vs_create(WAX_VS_MAX_DIMENSIONS + 1, WAX_VS_COSINE, NULL, 0, &e) != WAX_VS_ERR_CAPACITY) return 3;
    if (wax_vs_device_count(NULL) != WAX_VS_ERR_NULL) return 4;
    (void)wax_vs_device_count(&n);
    if (sizeof(wax_vs_candidate) != 24) return 5;
    wax_vs_destroy(NULL);
    printf("%s devices=%d\n", wax_vs_version(), (int)n);
    return 0;
}
