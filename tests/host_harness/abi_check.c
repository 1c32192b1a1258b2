/* Test infrastructure: include/prt.h is valid C99 and libprt.so links from plain C.  Calls only the entry points that
 * need no GPU; argument errors of the compute calls are checked as well (they must be rejected before any CUDA call). */
#include <stdio.h>
#include <string.h>

#include "prt.h"

int main(void)
{
    if (prt_abi_version() != PRT_ABI_VERSION) return 1;
    if (prt_device_count() < 0) return 2;
    if (prt_refine_workspace_bytes(1000) <= 0 || prt_bracket_workspace_bytes(100000, 64) <= 0) return 3;
    if (prt_iw_workspace_bytes(4, 200, 8) <= 0) return 4;
    /* null environment: PRT_ERR_ARG and a message, never a crash */
    if (prt_shoot_fan(NULL, 4, NULL, NULL, 0.0, 0.0, 1.0, 1e-9, 1e-6, 1, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL,
                      0, NULL) != PRT_ERR_ARG)
        return 5;
    if (strstr(prt_last_error(), "null environment") == NULL) return 6;
    if (prt_refine_brackets(NULL, 1, NULL, NULL, NULL, NULL, NULL, 0.0, 0.0, 1.0, 1e-9, 1e-6, 1, 1.0, 20, NULL, NULL, NULL, NULL,
                            NULL, NULL, NULL, NULL, NULL, 0, NULL) != PRT_ERR_ARG)
        return 7;
    if (prt_flat_earth(0, 1, 2, NULL, NULL, NULL, 0, NULL, NULL) != PRT_ERR_ARG) return 8;
    if (prt_env_destroy(NULL) != PRT_SUCCESS) return 9;
    printf("abi %d ok, %d CUDA device(s)\n", prt_abi_version(), prt_device_count());
    return 0;
}
