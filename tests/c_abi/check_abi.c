/* Plain C99 client of include/kvpress_b200.h: proves the header is valid C (no C++-isms), that the library links
 * from C with nothing but its own exports, and exercises the device-free entry points. Built and run by
 * tests/test_abi_symbols.py::test_header_is_valid_c_and_links_from_c. */
#include <stdio.h>
#include <string.h>

#include "kvpress_b200.h"

int main(void) {
    kvp_problem p;
    size_t bytes = 0, bytes_ea = 0;
    int launches = 0;
    memset(&p, 0, sizeof p);
    p.B = 1; p.Hkv = 8; p.Hq = 32; p.S = 131072; p.D = 128; p.n_kept = 39321; p.dtype = KVP_BF16;
    if (kvp_abi_version() != KVP_ABI_VERSION) return 1;
    if (strcmp(kvp_status_string(KVP_OK), "ok") != 0) return 2;
    if (kvp_workspace_bytes(&p, KVP_SCORER_KNORM, &bytes) != KVP_OK || bytes == 0) return 3;
    if (kvp_workspace_bytes(&p, KVP_SCORER_EXPECTED_ATTENTION, &bytes_ea) != KVP_OK || bytes_ea <= bytes) return 4;
    if (kvp_launches_per_compress(&p, KVP_SCORER_STREAMING, &launches) != KVP_OK || launches != 1) return 5;
    p.D = 12;
    if (kvp_workspace_bytes(&p, KVP_SCORER_KNORM, &bytes) != KVP_ERR_UNSUPPORTED_SHAPE) return 6;
    /* a compress call with a NULL cache pointer is refused before any CUDA call */
    p.D = 128;
    p.k_stride[0] = p.v_stride[0] = (int64_t)8 * 131072 * 128;
    p.k_stride[1] = p.v_stride[1] = (int64_t)131072 * 128;
    p.k_stride[2] = p.v_stride[2] = 128;
    if (kvp_knorm_compress(&p, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) != KVP_ERR_NULL_POINTER) return 7;
    printf("abi %d knorm_ws %zu ea_ws %zu\n", kvp_abi_version(), bytes, bytes_ea);
    return 0;
}
