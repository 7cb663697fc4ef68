/* Plain C99 caller of include/mrp_fg.h (the drop-in boundary): what a cgo / JNI / FFI binding would compile against.
 * Exercises the entry points that need no GPU and the error convention of the one that does. */
#include <stdio.h>
#include <string.h>

#include "mrp_fg.h"

int main(void) {
    mrp_fg_params p;
    int64_t ro[3] = {0, 46600, 93200}, seg = 0, chunks[64];
    int32_t tile = 0;
    mrp_fg_ctx* ctx = NULL;
    int n, i, rc;
    int64_t sum = 0;

    if (mrp_fg_abi_version() != MRP_FG_ABI_VERSION) return 10;
    if (mrp_fg_default_params(&p) != MRP_FG_OK || p.depth_levels != 10 || p.norm_window != 30000) return 11;
    if (mrp_fg_halo_rows_for(&p) != 30000 + 384 + 32 || mrp_fg_halo_rows_for(NULL) != 30416) return 12;
    if (mrp_fg_plan_shape(&p, 2, ro, 148, &seg, &tile) != MRP_FG_OK || seg < 2048 || seg > 16384 || p.norm_window % tile) return 13;
    ro[2] = 10;   /* decreasing offsets: an error code and a message, no exception */
    if (mrp_fg_plan_shape(&p, 2, ro, 148, &seg, &tile) != MRP_FG_ERR_INVALID || !strstr(mrp_fg_last_error(), "row_offsets")) return 14;
    n = mrp_fg_plan_host_chunks(50000000, 0, 1, 1, chunks, 64);
    if (n <= 3 || n > 64) return 15;
    for (i = 0; i < n; ++i) sum += chunks[i];
    if (sum != 50000000 || chunks[0] * 8 != chunks[3]) return 16;
    rc = mrp_fg_create(&p, 0, &ctx);
    if (rc == MRP_FG_OK) {
        if (mrp_fg_halo_rows(ctx) != 30416) return 17;
        mrp_fg_destroy(ctx);
        printf("abi_smoke ok (GPU present)\n");
    } else {
        if (rc != MRP_FG_ERR_CUDA || ctx != NULL || !strlen(mrp_fg_last_error())) return 18;   /* no CPU fallback: a clean refusal */
        printf("abi_smoke ok (no GPU: %s)\n", mrp_fg_last_error());
    }
    return 0;
}
