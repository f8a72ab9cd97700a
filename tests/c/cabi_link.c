/* Plain-C consumer of the drop-in boundary: include/ezflow_b200.h must compile as C99 and the shared library must link
 * and answer host-only entry points without Python or a GPU (tests/test_cabi_cpu.py builds and runs this). */
#include <stdio.h>
#include <string.h>

#include "ezflow_b200.h"

int main(void) {
  int h[4], w[4], sub_base[5], n_tiles = 0, n_groups = 0;
  int64_t group_bytes = 0, pair_bytes = 0;
  size_t ws = 0;
  char err[256];
  if (ezb_version() != 100) return 1;
  /* 1080p RAFT pyramid: floor pooling 135x240 -> 67x120 -> 33x60 -> 16x30 */
  if (ezb_corr_pyramid_layout(135, 240, 4, EZB_F16, h, w, sub_base, &n_tiles, &n_groups, &group_bytes, &pair_bytes) != 0) return 2;
  if (h[3] != 16 || w[3] != 30 || pair_bytes != (int64_t)n_tiles * n_groups * group_bytes) return 3;
  if (ezb_corr_pyramid_workspace_bytes(8, 256, 135, 240, 4, &ws) != 0 || ws == 0) return 4;
  if (ezb_convex_upsample_bwd_workspace_bytes(4, 2, 135, 240, &ws) != 0 || ws != (size_t)4 * 2 * 9 * 135 * 240 * 4) return 5;
  /* invalid arguments come back as a status and a message, never as an abort */
  if (ezb_corr_pyramid_layout(0, 4, 1, EZB_F32, 0, 0, 0, 0, 0, 0, 0) != EZB_ERR_INVALID) return 6;
  ezb_last_error(err, sizeof err);
  if (strstr(err, "bad pyramid shape") == 0) return 7;
  if (ezb_local_corr_fwd(0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 1, 1, 1.0f, 1.0f, 0, 0, 0, 0, 0) != EZB_ERR_INVALID) return 8;
  printf("cabi_link ok: %d tiles x %d groups x %lld bytes per pair\n", n_tiles, n_groups, (long long)group_bytes);
  return 0;
}
