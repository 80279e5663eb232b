/* Minimal C host of the drop-in boundary: links libgsrast_b200.so directly (no Python, no torch) and exercises the
 * entry points that need no GPU: ABI version, workspace sizing, layout introspection and the argument checks.
 * Built and run by tests/test_cabi.py::test_c_client_links_and_runs. */
#include <stdio.h>
#include <string.h>

#include "gsloss_b200.h"
#include "gsrast_b200.h"

int main(void)
{
    const int P = 1000000, W = 640, H = 480;
    gsr_geom_layout gl;
    gsr_img_layout il;
    gsr_binning_layout bl;
    gsr_gaussians g;
    gsr_camera cam;
    int rc;

    if (gsr_abi_version() != GSR_ABI_VERSION) return 1;
    if (gsr_geom_ws_bytes(P, W, H) == 0 || gsr_img_ws_bytes(W, H) == 0 || gsr_binning_ws_bytes(4000000) == 0) return 2;
    gsr_geom_layout_of(P, W, H, &gl);
    gsr_img_layout_of(W, H, &il);
    gsr_binning_layout_of(4000000, &bl);
    if (gl.total != gsr_geom_ws_bytes(P, W, H) || gl.rec >= gl.total || gl.counters >= gl.total) return 3;
    if (il.total != gsr_img_ws_bytes(W, H) || il.ranges >= il.total) return 4;
    if (bl.total != gsr_binning_ws_bytes(4000000) || bl.point_list >= bl.total) return 5;
    if (gsr_slam_loss_ws_bytes(W, H) < (size_t)9 * W * H * sizeof(float)) return 6;

    /* argument checks answer before any CUDA call */
    memset(&g, 0, sizeof g);
    memset(&cam, 0, sizeof cam);
    g.P = 10;
    cam.width = W;
    cam.height = H;
    rc = gsr_forward_preprocess(NULL, &g, &cam, NULL, NULL, 0, NULL, 0, NULL);
    if (rc != GSR_ERR_INVALID || gsr_last_error() == NULL || strlen(gsr_last_error()) == 0) return 7;
    if (gsr_adam_step(NULL, NULL, NULL, NULL, NULL, 8, 1, NULL, NULL, 0.9, 0.999, 1e-15, 0, 1.0f, 0) != GSR_ERR_INVALID) return 8;
    printf("abi %d geom %zu img %zu binning(4M) %zu loss %zu stages %d\n", gsr_abi_version(), gl.total, il.total, bl.total,
           gsr_slam_loss_ws_bytes(W, H), gsr_profile_num_stages());
    return 0;
}
