/* abi_c.c - include/cuboid_cuda.h is a C header: this file is compiled as C99 and linked against libcuboid_cuda.so. */
#include <stdio.h>
#include <string.h>

#include "cuboid_cuda.h"

int main(void) {
    cuboid_params p;
    cuboid_frame_result r;
    double H[16], pose[7];
    const float T[16] = {1, 0, 0, 0.5f, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    cuboid_default_params(&p);
    memset(&r, 0, sizeof r);
    if (cuboid_params_size() != (int)sizeof p || cuboid_frame_result_size() != (int)sizeof r) return 1;
    cuboid_pose_from_transform(T, H, pose);
    if (pose[0] != -0.5) return 2;
    printf("%s\n", cuboid_strerror(CUBOID_E_NO_DEVICE));
    return 0;
}
