/* A plain-C client of the drop-in boundary: includes include/navgym_b200.h as C99, links
 * libnavgym_b200.so and checks the ABI the way a non-Python host (cgo / JNI / FFI) would see it.
 * Built and run by tests/test_abi.py; it makes no CUDA call that needs a device. */
#include <stdio.h>
#include <string.h>
#include "navgym_b200.h"

int main(void)
{
    navgym_step_args_t a;
    memset(&a, 0, sizeof a);
    if (navgym_abi_version() != 1) return 1;
    if (navgym_sizeof_step_args() != (int)sizeof(navgym_step_args_t)) return 2;
    if (navgym_sizeof_map() != (int)sizeof(navgym_map_t)) return 3;
    if (navgym_sizeof_her_args() != (int)sizeof(navgym_her_args_t)) return 4;
    if (navgym_sizeof_peds_args() != (int)sizeof(navgym_peds_args_t)) return 5;
    if (navgym_sizeof_scan_args() != (int)sizeof(navgym_scan_args_t)) return 6;
    if (navgym_sizeof_plan_args() != (int)sizeof(navgym_plan_args_t)) return 7;
    if (navgym_sizeof_plan_map() != (int)sizeof(navgym_plan_map_t)) return 8;
    if (navgym_sizeof_move_args() != (int)sizeof(navgym_move_args_t)) return 9;
    /* an empty batch is a no-op, not an error, and touches no device */
    if (navgym_step_batch(&a, NULL) != 0) return 10;
    if (navgym_reset_obs_batch(&a, NULL) != 0) return 11;
    if (!navgym_error_string(0)) return 12;
    /* host-only helper: 4-connected BFS over a tiny grid */
    {
        const uint8_t blocked[9] = {0, 0, 0, 1, 1, 0, 0, 0, 0};
        int32_t dist[9];
        navgym_grid_bfs(blocked, 3, 3, 0, 0, dist);
        if (dist[0] != 0 || dist[2] != 2 || dist[5] != 3 || dist[6] != 6 || dist[3] != -1) return 13;
    }
    printf("abi ok: version %d, step args %d bytes, devices %d\n", navgym_abi_version(),
           navgym_sizeof_step_args(), navgym_device_count());
    return 0;
}
