/*
 * synth.c — seeded synthetic D435-shaped depth frames (host-only tool).
 *
 * The reference ships no bag files (/root/reference/.gitignore:1), so every
 * test and bench input is produced here: a ray-cast table plane plus up to 16 cuboids, depth
 * noise N(0, sigma), quantisation to uint16 millimetres, and a fraction of invalid (0) pixels.
 * The intrinsics default to the D435 depth K recorded at /root/reference/README.md:78.
 *
 * Counter-based RNG (splitmix64 of seed/pixel/stream) so a frame does not depend on thread
 * count or generation order. This file is NOT part of the product path (libcuboid_cuda never
 * links it) and NOT part of the oracle; it only makes inputs.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
    double L, W, H;      /* cuboid extents (m): x,y,z in its own frame, centred */
    double px, py;       /* centre on the table, world frame (m) */
    double yaw;          /* rotation about world Z (rad) */
} synth_box;

typedef struct {
    int32_t w, h;
    double fx, fy, cx, cy;
    double cam_h;        /* camera height above the table (m) */
    double tilt;         /* optical axis below horizontal (rad) */
    double noise_sigma;  /* metres, applied to z before mm quantisation */
    double invalid_frac; /* fraction of pixels forced to 0 */
    uint64_t seed;
    int32_t n_boxes;
    int32_t _pad;
    synth_box box[16];
} synth_scene;

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline double u01(uint64_t seed, uint64_t pix, uint64_t stream) {
    uint64_t r = splitmix64(splitmix64(seed * 0x100000001B3ull + stream) ^ (pix * 0xD6E8FEB86659FD93ull));
    return ((double)(r >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

/* world frame: Z up, table = {Z = 0}; camera at (0,0,cam_h) looking along +X, pitched down by tilt.
 * camera axes in world: x_c = (0,-1,0), y_c = (-sin t, 0, -cos t), z_c = (cos t, 0, -sin t). */
void synth_depth_frame(const synth_scene* sc, uint16_t* depth) {
    const double st = sin(sc->tilt), ct = cos(sc->tilt);
    for (int v = 0; v < sc->h; ++v) {
        for (int u = 0; u < sc->w; ++u) {
            const uint64_t pix = (uint64_t)v * (uint64_t)sc->w + (uint64_t)u;
            /* ray in camera frame with z component 1, so t == camera-frame depth z */
            const double dxc = ((double)u - sc->cx) / sc->fx;
            const double dyc = ((double)v - sc->cy) / sc->fy;
            /* ray direction in world */
            const double dwx = -st * dyc + ct;
            const double dwy = -dxc;
            const double dwz = -ct * dyc - st;
            double best = INFINITY;
            if (dwz < -1e-12) {
                const double t = sc->cam_h / (-dwz);
                if (t > 0) best = t;
            }
            for (int b = 0; b < sc->n_boxes; ++b) {
                const synth_box* bx = &sc->box[b];
                const double cy_ = cos(bx->yaw), sy_ = sin(bx->yaw);
                /* origin and direction in the box frame (box centre at (px,py,H/2)) */
                const double ox = -bx->px, oy = -bx->py, oz = sc->cam_h - 0.5 * bx->H;
                const double o[3] = { cy_ * ox + sy_ * oy, -sy_ * ox + cy_ * oy, oz };
                const double d[3] = { cy_ * dwx + sy_ * dwy, -sy_ * dwx + cy_ * dwy, dwz };
                const double hs[3] = { 0.5 * bx->L, 0.5 * bx->W, 0.5 * bx->H };
                double tn = -INFINITY, tf = INFINITY;
                int ok = 1;
                for (int a = 0; a < 3; ++a) {
                    if (fabs(d[a]) < 1e-15) {
                        if (o[a] < -hs[a] || o[a] > hs[a]) { ok = 0; break; }
                    } else {
                        double t0 = (-hs[a] - o[a]) / d[a], t1 = (hs[a] - o[a]) / d[a];
                        if (t0 > t1) { double s = t0; t0 = t1; t1 = s; }
                        if (t0 > tn) tn = t0;
                        if (t1 < tf) tf = t1;
                        if (tn > tf) { ok = 0; break; }
                    }
                }
                if (ok && tn > 0 && tn < best) best = tn;
            }
            uint16_t out = 0;
            if (isfinite(best)) {
                /* ~N(0,1): Irwin-Hall sum of four 16-bit uniforms from one hash (variance 4/12), no libm calls */
                const uint64_t r = splitmix64(splitmix64(sc->seed * 0x100000001B3ull + 1) ^ (pix * 0xD6E8FEB86659FD93ull));
                const double s4 = (double)(r & 0xffff) + (double)((r >> 16) & 0xffff) + (double)((r >> 32) & 0xffff) + (double)(r >> 48);
                const double g = (s4 * (1.0 / 65536.0) - 1.999969482421875) * 1.7320508075688772;
                const double z = best + sc->noise_sigma * g;
                const double mm = floor(z * 1000.0 + 0.5);
                if (mm >= 1.0 && mm <= 65535.0) out = (uint16_t)mm;
            }
            if (u01(sc->seed, pix, 3) < sc->invalid_frac) out = 0;
            depth[pix] = out;
        }
    }
}

int synth_scene_size(void) { return (int)sizeof(synth_scene); }
