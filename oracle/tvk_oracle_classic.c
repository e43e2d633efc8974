/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's CLASSIC per-brick ray-caster (SURVEY 8a rows
 * R12): what GLRaycaster draws for one sub-frame, pixel by pixel.
 *
 * Follows, under /root/reference/Tuvok/:
 *   Renderer/GL/GLRaycaster.cpp:383-479   Render3DPreLoop / Render3DInLoop (near-plane pass, front faces into the
 *                                         RGBA16F entry buffer :97, back faces with the ray-cast shader)
 *   Renderer/GL/GLRaycaster.cpp:239-300   SetBrickDepShaderVars (fRayStepsize, fStepScale, vVoxelStepsize)
 *   Renderer/GL/GLRaycaster.cpp:589-612   ComputeEyeToTextureMatrix (computed by the caller, see oracle/host.py)
 *   Shaders/GLRaycaster-1D-FS.glsl:51-82, GLRaycaster-1D-light-FS.glsl:83-134   main()
 *   Shaders/GLRaycaster-2D-FS.glsl:52-94, GLRaycaster-2D-light-FS.glsl:62-117   main() (2D transfer function)
 *   Shaders/GLRaycaster-ISO-FS.glsl:56-104, RefineIsosurface.glsl:37-53         first hit per brick, refined; the
 *                                         hit survives if it passes the GL_LESS depth test (GLRenderer.cpp:139) against
 *                                         the buffer cleared to 1.0 (GLRaycaster.cpp:391-394); depth compared in fp32
 *   Shaders/Compose-FS.glsl:49-75, GLRenderer::ComposeSurfaceImage (GLRenderer.cpp:2763-2811)   isosurface shading
 *   Shaders/VRender1D.glsl, VRender1DLit.glsl, Volume3D.glsl (sampleVolume, ComputeGradient, ComputeNormal),
 *   Shaders/lighting.glsl, Shaders/Compositing.glsl:33-38
 *   Renderer/GL/GLRenderer.cpp:152-153    frame-buffer blend ONE_MINUS_DST_ALPHA / ONE between bricks
 * Parity status: "parity unpinned" like the ray-guided restatement (the reference's pixels need an OpenGL driver).
 * Restatement choices: brick faces are intersected analytically with the pixel-centre ray instead of being
 * rasterised; a brick whose front face lies behind the near plane starts at the near plane (the reference reads
 * whatever the previous brick left in the entry buffer there); the entry point is rounded to half precision as the
 * RGBA16F buffer does, the exit point is not (it is an interpolated varying).
 * Floating-point contract as in tvk_oracle_render.c.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "tvk_cuda.h"

static float half_round(float f) {          /* float -> IEEE binary16 (round to nearest even) -> float */
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = x & 0x80000000u;
  x &= 0x7FFFFFFFu;
  if (x >= 0x7F800000u) return f;                              /* inf / nan */
  if (x >= 0x477FF000u) { x = 0x7F800000u | sign; memcpy(&f, &x, 4); return f; }   /* rounds to >= 65520 -> inf */
  float a;
  memcpy(&a, &x, 4);
  if (x < 0x38800000u) {                                       /* below 2^-14: half subnormal, quantum 2^-24 */
    const float q = 5.9604644775390625e-08f;
    a = rintf(a / q) * q;
  } else {
    const uint32_t rem = x & 0x1FFFu, base = x & ~0x1FFFu;     /* keep 10 mantissa bits */
    uint32_t r = base;
    if (rem > 0x1000u || (rem == 0x1000u && (base & 0x2000u))) r += 0x2000u;
    memcpy(&a, &r, 4);
  }
  uint32_t y;
  memcpy(&y, &a, 4);
  y |= sign;
  memcpy(&a, &y, 4);
  return a;
}

static inline float dot3c(float ax, float ay, float az, float bx, float by, float bz) { return fmaf(az, bz, fmaf(ay, by, ax * bx)); }
static inline float clamp01c(float v) { return fminf(fmaxf(v, 0.0f), 1.0f); }
static void xformc(const float* m, float x, float y, float z, float w, float o[4]) {
  o[0] = fmaf(w, m[12], fmaf(z, m[8], fmaf(y, m[4], x * m[0])));
  o[1] = fmaf(w, m[13], fmaf(z, m[9], fmaf(y, m[5], x * m[1])));
  o[2] = fmaf(w, m[14], fmaf(z, m[10], fmaf(y, m[6], x * m[2])));
  o[3] = fmaf(w, m[15], fmaf(z, m[11], fmaf(y, m[7], x * m[3])));
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int f2i_sat(float f) {          /* (int)f with the saturation CUDA's conversion has; NaN -> 0 */
  if (!(f == f)) return 0;
  if (f >= 2147483648.0f) return 2147483647;
  if (f <= -2147483648.0f) return -2147483647 - 1;
  return (int)f;
}

static float brick_sample(const uint8_t* data, uint32_t bpv, const uint32_t vox[3], float tx, float ty, float tz) {
  const float u = fmaf(tx, (float)vox[0], -0.5f), v = fmaf(ty, (float)vox[1], -0.5f), w = fmaf(tz, (float)vox[2], -0.5f);
  const float fu = floorf(u), fv = floorf(v), fw = floorf(w);
  const float a = u - fu, b = v - fv, g = w - fw;
  const int mx = (int)vox[0] - 1, my = (int)vox[1] - 1, mz = (int)vox[2] - 1;
  const int x0 = clampi(f2i_sat(fu), 0, mx), x1 = clampi(f2i_sat(fu) == 2147483647 ? 2147483647 : f2i_sat(fu) + 1, 0, mx);
  const int y0 = clampi(f2i_sat(fv), 0, my), y1 = clampi(f2i_sat(fv) == 2147483647 ? 2147483647 : f2i_sat(fv) + 1, 0, my);
  const int z0 = clampi(f2i_sat(fw), 0, mz), z1 = clampi(f2i_sat(fw) == 2147483647 ? 2147483647 : f2i_sat(fw) + 1, 0, mz);
  const size_t sx = vox[0], sxy = (size_t)vox[0] * vox[1];
#define VOX(X, Y, Z) (bpv == 1 ? (float)data[(X) + (Y) * sx + (Z) * sxy] : (float)((const uint16_t*)data)[(X) + (Y) * sx + (Z) * sxy])
  const float v000 = VOX(x0, y0, z0), v100 = VOX(x1, y0, z0), v010 = VOX(x0, y1, z0), v110 = VOX(x1, y1, z0);
  const float v001 = VOX(x0, y0, z1), v101 = VOX(x1, y0, z1), v011 = VOX(x0, y1, z1), v111 = VOX(x1, y1, z1);
#undef VOX
  const float c00 = fmaf(a, v100 - v000, v000), c10 = fmaf(a, v110 - v010, v010);
  const float c01 = fmaf(a, v101 - v001, v001), c11 = fmaf(a, v111 - v011, v011);
  const float c0 = fmaf(b, c10 - c00, c00), c1 = fmaf(b, c11 - c01, c01);
  return fmaf(g, c1 - c0, c0) * (bpv == 1 ? 1.0f / 255.0f : 1.0f / 65535.0f);
}

static void brick_gradient(const uint8_t* data, uint32_t bpv, const uint32_t vox[3], float tx, float ty, float tz, float g[3]) {
  const float hx = 1.0f / (float)vox[0], hy = 1.0f / (float)vox[1], hz = 1.0f / (float)vox[2];
  const float gxp = brick_sample(data, bpv, vox, tx + hx, ty, tz), gxm = brick_sample(data, bpv, vox, tx - hx, ty, tz);
  const float gyp = brick_sample(data, bpv, vox, tx, ty - hy, tz), gym = brick_sample(data, bpv, vox, tx, ty + hy, tz);
  const float gzp = brick_sample(data, bpv, vox, tx, ty, tz + hz), gzm = brick_sample(data, bpv, vox, tx, ty, tz - hz);
  g[0] = (gxm - gxp) / 2.0f; g[1] = (gyp - gym) / 2.0f; g[2] = (gzm - gzp) / 2.0f;
}

static void classic_normal(const tvk_classic_params* p, const float g[3], float N[3]) {
  const float gx = g[0] * p->domain_scale[0], gy = g[1] * p->domain_scale[1], gz = g[2] * p->domain_scale[2];
  N[0] = fmaf(gz, p->normal_matrix[6], fmaf(gy, p->normal_matrix[3], gx * p->normal_matrix[0]));
  N[1] = fmaf(gz, p->normal_matrix[7], fmaf(gy, p->normal_matrix[4], gx * p->normal_matrix[1]));
  N[2] = fmaf(gz, p->normal_matrix[8], fmaf(gy, p->normal_matrix[5], gx * p->normal_matrix[2]));
  const float nl = sqrtf(dot3c(N[0], N[1], N[2], N[0], N[1], N[2]));
  if (nl > 0.0f) { N[0] /= nl; N[1] /= nl; N[2] /= nl; }
}

static inline int nearest_c(float s, uint32_t n) { return clampi(f2i_sat(s * (float)n), 0, (int)n - 1); }

static void classic_compose(const tvk_classic_params* p, const float hit[4], const float hn[3], float out[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.0f;
  if (hit[3] == 0.0f) return;
  const float nx = hn[0], ny = hn[1], nz = fabsf(hn[2]);
  float vx = 0.0f - hit[0], vy = 0.0f - hit[1], vz = 0.0f - hit[2];
  float l = sqrtf(dot3c(vx, vy, vz, vx, vy, vz));
  vx /= l; vy /= l; vz /= l;
  const float d2 = 2.0f * dot3c(nx, ny, nz, vx, vy, vz);
  float rx = fmaf(-d2, nx, vx), ry = fmaf(-d2, ny, vy), rz = fmaf(-d2, nz, vz);
  l = sqrtf(dot3c(rx, ry, rz, rx, ry, rz));
  rx /= l; ry /= l; rz /= l;
  const float ndl = fmaxf(fabsf(dot3c(nx, ny, nz, -p->light_dir[0], -p->light_dir[1], -p->light_dir[2])), 0.0f);
  float sp = fmaxf(dot3c(rx, ry, rz, p->light_dir[0], p->light_dir[1], p->light_dir[2]), 0.0f);
  sp *= sp; sp *= sp; sp *= sp;
  for (int k = 0; k < 3; ++k)
    out[k] = clamp01c(fmaf(p->light_specular[k], sp, fmaf(p->light_diffuse[k] * p->iso_color[k], ndl, p->light_ambient[k])));
  out[3] = 1.0f;
}

/* tf_rgba8: [tf_h][tf_w][4]; tf_h == 1 for the 1D modes */
int orc_raycast_classic(const tvk_classic_params* p, const tvk_classic_brick* bricks, uint32_t n, const uint8_t* store,
                        const uint8_t* tf_rgba8, uint32_t tf_n, uint32_t tf_h, float* image, uint64_t* samples_out) {
  if (!p || (n && !bricks) || !store || !tf_rgba8 || !image || !tf_h) return 1;
  const int lit = p->lighting && p->render_mode != TVK_RM_ISOSURFACE;
  uint64_t total = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : total)
  for (int64_t py = 0; py < (int64_t)p->height; ++py) {
    for (uint32_t px = 0; px < p->width; ++px) {
      const float nx = fmaf(((float)px + 0.5f) / (float)p->width, 2.0f, -1.0f), ny = fmaf(((float)py + 0.5f) / (float)p->height, 2.0f, -1.0f);
      float n4[4], O[4], D[4];
      xformc(p->inv_projection, nx, ny, -1.0f, 1.0f, n4);
      const float ex = n4[0] / n4[3], ey = n4[1] / n4[3], ez = n4[2] / n4[3];
      xformc(p->inv_model_view, 0.f, 0.f, 0.f, 1.0f, O);
      xformc(p->inv_model_view, ex, ey, ez, 0.0f, D);
      float dst[4] = {0.f, 0.f, 0.f, 0.f};
      float hit[4] = {0.f, 0.f, 0.f, 0.f}, hn[3] = {0.f, 0.f, 0.f}, best_depth = 1.0f;
      for (uint32_t bi = 0; bi < n; ++bi) {
        const tvk_classic_brick* b = &bricks[bi];
        float s0 = -INFINITY, s1 = INFINITY;
        int miss = 0;
        for (int k = 0; k < 3; ++k) {
          const float lo = b->center[k] - b->extension[k] / 2.0f, hi = b->center[k] + b->extension[k] / 2.0f;
          if (D[k] == 0.0f) { if (O[k] < lo || O[k] > hi) miss = 1; continue; }
          const float inv = 1.0f / D[k];
          float ta = (lo - O[k]) * inv, tb = (hi - O[k]) * inv;
          if (ta > tb) { const float t = ta; ta = tb; tb = t; }
          s0 = fmaxf(s0, ta); s1 = fminf(s1, tb);
        }
        if (p->clip_enabled && !miss) {       /* RenderBox + Clipper::BoxPlane (GLRaycaster.cpp:327-337): n.p + d <= 0 is kept */
          const float dn = dot3c(p->clip_plane[0], p->clip_plane[1], p->clip_plane[2], D[0], D[1], D[2]);
          const float on = dot3c(p->clip_plane[0], p->clip_plane[1], p->clip_plane[2], O[0], O[1], O[2]) + p->clip_plane[3];
          if (dn == 0.0f) { if (on > 0.0f) miss = 1; }
          else {
            const float t = -on / dn;
            if (dn > 0.0f) s1 = fminf(s1, t); else s0 = fmaxf(s0, t);
          }
        }
        const float se = fmaxf(s0, 1.0f);
        if (miss || !(s1 > se)) continue;
        const float enx = half_round(ex * se), eny = half_round(ey * se), enz = half_round(ez * se);
        const float xx = ex * s1, xy = ey * s1, xz = ez * s1;
        float te[4], tx[4];
        xformc(b->eye_to_tex, enx, eny, enz, 1.0f, te);
        xformc(b->eye_to_tex, xx, xy, xz, 1.0f, tx);
        float dx = xx - enx, dy = xy - eny, dz = xz - enz;
        const float len = sqrtf(dot3c(dx, dy, dz, dx, dy, dz));
        const int steps = f2i_sat(len / b->ray_step) + 1;
        const float q = len / b->ray_step;
        const float itx = (tx[0] - te[0]) / q, ity = (tx[1] - te[1]) / q, itz = (tx[2] - te[2]) / q;
        dx /= len; dy /= len; dz /= len;
        const float ix = dx * b->ray_step, iy = dy * b->ray_step, iz = dz * b->ray_step;
        const uint8_t* data = store + b->store_offset;
        float col[4] = {0.f, 0.f, 0.f, 0.f};
        float ptx = te[0], pty = te[1], ptz = te[2], pex = enx, pey = eny, pez = enz;
        const uint32_t bpv = p->bytes_per_voxel;
        if (p->render_mode == TVK_RM_ISOSURFACE) {
          int found = 0;
          for (int i = 0; i < steps; ++i) {
            const float v = brick_sample(data, bpv, b->voxels, ptx, pty, ptz);
            total++;
            if (v >= p->isoval) { found = 1; break; }
            ptx += itx; pty += ity; ptz += itz;
          }
          if (!found) continue;
          float rx = itx / 2.0f, ry = ity / 2.0f, rz = itz / 2.0f;
          ptx -= rx; pty -= ry; ptz -= rz;
          for (int i = 0; i < 5; ++i) {
            rx /= 2.0f; ry /= 2.0f; rz /= 2.0f;
            const float v = brick_sample(data, bpv, b->voxels, ptx, pty, ptz);
            if (v >= p->isoval) { ptx -= rx; pty -= ry; ptz -= rz; } else { ptx += rx; pty += ry; ptz += rz; }
          }
          const float ltx = tx[0] - te[0], lty = tx[1] - te[1], ltz = tx[2] - te[2];
          const float len_tex = sqrtf(dot3c(ltx, lty, ltz, ltx, lty, ltz));
          const float hx = ptx - te[0], hy = pty - te[1], hz = ptz - te[2];
          const float f = sqrtf(dot3c(hx, hy, hz, hx, hy, hz)) / len_tex;
          const float hpx = fmaf(xx, f, enx * (1.0f - f)), hpy = fmaf(xy, f, eny * (1.0f - f)), hpz = fmaf(xz, f, enz * (1.0f - f));
          const float depth = clamp01c(p->proj_param[0] + p->proj_param[1] / (0.0f - hpz));
          if (depth < best_depth) {
            best_depth = depth;
            hit[0] = hpx; hit[1] = hpy; hit[2] = hpz; hit[3] = f;
            float g[3];
            brick_gradient(data, bpv, b->voxels, ptx, pty, ptz, g);
            classic_normal(p, g, hn);
          }
          continue;
        }
        for (int i = 0; i < steps; ++i) {
          const float v = brick_sample(data, bpv, b->voxels, ptx, pty, ptz);
          float c[4], g[3];
          const int two_d = p->render_mode == TVK_RM_2DTRANS;
          if (two_d || lit) brick_gradient(data, bpv, b->voxels, ptx, pty, ptz, g);
          size_t ti;
          if (two_d) {
            const float gm = sqrtf(dot3c(g[0], g[1], g[2], g[0], g[1], g[2]));
            ti = (size_t)nearest_c(fmaf(-gm, p->gradient_scale, 1.0f), tf_h) * tf_n + nearest_c(v * p->trans_scale, tf_n);
          } else ti = (size_t)clampi(f2i_sat(v * p->trans_scale * (float)tf_n), 0, (int)tf_n - 1);
          for (int k = 0; k < 4; ++k) c[k] = (float)tf_rgba8[4 * ti + k] * (1.0f / 255.0f);
          total++;
          if (lit) {
            float N[3];
            classic_normal(p, g, N);
            const float Nx = N[0], Ny = N[1], Nz = N[2];
            float vx = 0.0f - pex, vy = 0.0f - pey, vz = 0.0f - pez;
            float l = sqrtf(dot3c(vx, vy, vz, vx, vy, vz));
            vx /= l; vy /= l; vz /= l;
            const float d2 = 2.0f * dot3c(Nx, Ny, Nz, vx, vy, vz);
            float rx = fmaf(-d2, Nx, vx), ry = fmaf(-d2, Ny, vy), rz = fmaf(-d2, Nz, vz);
            l = sqrtf(dot3c(rx, ry, rz, rx, ry, rz));
            rx /= l; ry /= l; rz /= l;
            const float ndl = fmaxf(fabsf(dot3c(Nx, Ny, Nz, p->light_dir[0], p->light_dir[1], p->light_dir[2])), 0.0f);
            float sp = fmaxf(dot3c(rx, ry, rz, p->light_dir[0], p->light_dir[1], p->light_dir[2]), 0.0f);
            sp *= sp; sp *= sp; sp *= sp;
            c[0] = clamp01c(fmaf(p->light_specular[0], sp, fmaf(p->light_diffuse[0] * c[0], ndl, p->light_ambient[0])));
            c[1] = clamp01c(fmaf(p->light_specular[1], sp, fmaf(p->light_diffuse[1] * c[1], ndl, p->light_ambient[1])));
            c[2] = clamp01c(fmaf(p->light_specular[2], sp, fmaf(p->light_diffuse[2] * c[2], ndl, p->light_ambient[2])));
          }
          c[3] = (p->step_scale == 1.0f) ? 1.0f - (1.0f - c[3]) : 1.0f - powf(1.0f - c[3], p->step_scale);
          const float wgt = (1.0f - col[3]) * c[3];
          col[0] = fmaf(c[0], wgt, col[0]); col[1] = fmaf(c[1], wgt, col[1]); col[2] = fmaf(c[2], wgt, col[2]);
          col[3] = fmaf(c[3], 1.0f - col[3], col[3]);
          pex += ix; pey += iy; pez += iz;
          ptx += itx; pty += ity; ptz += itz;
          if (col[3] >= 0.99f) break;
        }
        const float k = 1.0f - dst[3];
        dst[0] = fmaf(col[0], k, dst[0]); dst[1] = fmaf(col[1], k, dst[1]); dst[2] = fmaf(col[2], k, dst[2]); dst[3] = fmaf(col[3], k, dst[3]);
      }
      if (p->render_mode == TVK_RM_ISOSURFACE) classic_compose(p, hit, hn, dst);
      float* o = image + 4 * ((size_t)py * p->width + px);
      o[0] = dst[0]; o[1] = dst[1]; o[2] = dst[2]; o[3] = dst[3];
    }
  }
  if (samples_out) *samples_out = total;
  return 0;
}
