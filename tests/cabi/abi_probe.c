/* TEST INFRASTRUCTURE: exercises the C-ABI of libmvpraymarch_b200.so from plain C, through include/mvpraymarch_b200.h
 * alone (no Python, no torch, no CUDA headers): dlopen, symbol lookup, struct layout, workspace query and every
 * argument-error path that is validated before any device work.  The interface it pins is the replacement for the
 * reference's pybind module (/root/reference/extensions/mvpraymarch/mvpraymarch.cpp:180-396).
 * usage: abi_probe /path/to/libmvpraymarch_b200.so     (prints one "key value" line per fact; exit code 0 = all ok) */
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include "mvpraymarch_b200.h"

#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!h) { printf("FAIL dlopen %s\n", dlerror()); return 1; }
    int (*abi)(void) = (int (*)(void))dlsym(h, "mvp_abi_version");
    const char *(*errstr)(int) = (const char *(*)(int))dlsym(h, "mvp_error_string");
    size_t (*wsbytes)(const mvp_shape *) = (size_t (*)(const mvp_shape *))dlsym(h, "mvp_workspace_bytes");
    int (*fwd)(const mvp_forward_args *, void *) = (int (*)(const mvp_forward_args *, void *))dlsym(h, "mvp_raymarch_forward");
    int (*bwd)(const mvp_backward_args *, void *) = (int (*)(const mvp_backward_args *, void *))dlsym(h, "mvp_raymarch_backward");
    int (*accel)(const mvp_shape *, uint32_t, const int32_t *, const float *, const float *, const float *, const float *, const float *,
                 void *, size_t, void *) = (int (*)(const mvp_shape *, uint32_t, const int32_t *, const float *, const float *,
                                                    const float *, const float *, const float *, void *, size_t, void *))dlsym(h, "mvp_build_accel");
    int (*accel_cam)(const mvp_shape *, uint32_t, const int32_t *, const mvp_camera *, const float *, const float *, const float *, void *, size_t,
                     void *) = (int (*)(const mvp_shape *, uint32_t, const int32_t *, const mvp_camera *, const float *, const float *,
                                        const float *, void *, size_t, void *))dlsym(h, "mvp_build_accel_camera");
    CHECK(abi && errstr && wsbytes && fwd && bwd && accel && accel_cam);
    CHECK(abi() == MVP_ABI_VERSION);
    printf("abi %d\nsizeof_shape %zu\nsizeof_forward_args %zu\nsizeof_backward_args %zu\nsizeof_camera %zu\n", abi(), sizeof(mvp_shape),
           sizeof(mvp_forward_args), sizeof(mvp_backward_args), sizeof(mvp_camera));

    mvp_shape c3 = {80, 1024, 667, 16384, 8, 8, 8}, bad = {1, 0, 8, 4, 2, 2, 2};
    size_t ws = wsbytes(&c3);
    CHECK(ws > 0 && ws % 256 == 0 && wsbytes(&bad) == 0 && wsbytes(NULL) == 0);
    printf("workspace_bytes_c3 %zu\n", ws);

    mvp_forward_args a;
    memset(&a, 0, sizeof a);
    a.shape = (mvp_shape){1, 8, 8, 4, 2, 2, 2};
    a.stepsize = 0.1f;
    CHECK(fwd(NULL, NULL) == MVP_ERR_NULL);
    CHECK(fwd(&a, NULL) == MVP_ERR_STRUCT);                      /* struct_size not set */
    a.struct_size = (uint32_t)sizeof a - (uint32_t)sizeof(mvp_camera);   /* the ABI-v7 length: a stale caller */
    CHECK(fwd(&a, NULL) == MVP_ERR_STRUCT);
    a.struct_size = (uint32_t)sizeof a;
    CHECK(fwd(&a, NULL) == MVP_ERR_NULL);                        /* required pointers missing */
    float *dummy = (float *)(uintptr_t)256;
    a.raypos = a.raydir = a.tminmax = a.primpos = a.primrot = a.primscale = a.tplate = dummy;
    a.rayrgba = dummy;
    a.workspace = dummy;
    a.workspace_bytes = 16;
    CHECK(fwd(&a, NULL) == MVP_ERR_WORKSPACE);
    a.workspace_bytes = (size_t)1 << 30;
    a.algo = 7;
    CHECK(fwd(&a, NULL) == MVP_ERR_ALGO);
    a.algo = 1;
    CHECK(fwd(&a, NULL) == MVP_ERR_NULL);                        /* algo 1 without a warp field */
    a.algo = 0;
    a.stepsize = 0.f;
    CHECK(fwd(&a, NULL) == MVP_ERR_STEPSIZE);
    a.stepsize = 0.1f;
    a.shape.K = 0;
    CHECK(fwd(&a, NULL) == MVP_ERR_SHAPE);
    a.shape.K = 4;
    a.tplate = (const float *)(uintptr_t)260;                    /* 4-byte aligned only */
    CHECK(fwd(&a, NULL) == MVP_ERR_ALIGN);
    a.tplate = dummy;
    a.rayrgb_nchw = dummy;                                        /* image-plane outputs come in pairs */
    CHECK(fwd(&a, NULL) == MVP_ERR_NULL);
    a.rayrgb_nchw = NULL;
    a.raysat = dummy;                                             /* raysat without rayaux */
    CHECK(fwd(&a, NULL) == MVP_ERR_NULL);
    a.raysat = NULL;
    /* rays from the camera (mvp_camera): the ray tensors may be NULL, the four camera arrays come together, volradius > 0 */
    a.raypos = a.raydir = a.tminmax = NULL;
    CHECK(fwd(&a, NULL) == MVP_ERR_NULL);
    a.camera.viewpos = dummy;
    CHECK(fwd(&a, NULL) == MVP_ERR_NULL);                        /* viewrot / focal / princpt missing */
    a.camera.viewrot = a.camera.focal = a.camera.princpt = dummy;
    CHECK(fwd(&a, NULL) == MVP_ERR_CAMERA);                      /* volradius = 0 */
    a.camera.volradius = 256.f;
    a.stepsize = 0.f;
    CHECK(fwd(&a, NULL) == MVP_ERR_STEPSIZE);                    /* i.e. the camera was accepted in place of the ray tensors */
    {
        mvp_camera cam0;
        memset(&cam0, 0, sizeof cam0);
        CHECK(accel_cam(&c3, 0, NULL, &cam0, dummy, dummy, dummy, dummy, (size_t)1 << 40, NULL) == MVP_ERR_NULL);
    }

    mvp_backward_args b;
    memset(&b, 0, sizeof b);
    CHECK(bwd(&b, NULL) == MVP_ERR_STRUCT);
    b.struct_size = (uint32_t)sizeof b;
    CHECK(bwd(&b, NULL) == MVP_ERR_NULL);
    CHECK(accel(&c3, 0, NULL, NULL, NULL, NULL, NULL, NULL, NULL, 0, NULL) == MVP_ERR_NULL);
    CHECK(strstr(errstr(MVP_ERR_WORKSPACE), "workspace") != NULL && strstr(errstr(MVP_ERR_STRUCT), "struct_size") != NULL);
    printf("ok 1\n");
    dlclose(h);
    return 0;
}
