// TEST INFRASTRUCTURE ONLY -- what the CPU harness binds to the reference's shader programs (see build_glslref.py).
#pragma once
#include "glsl_runtime.h"
#include "tvk_cuda.h"

struct Scene {
  const tvk_raycast_params* p = nullptr;
  glsl::sampler2D ray_start_point, ray_start_aux, hit_pos, hit_normal;
  glsl::sampler1D tf1;
  glsl::sampler2D tf2;
  glsl::usampler3D meta;
  glsl::sampler3D pool;
  glsl::uimage1D hash1;
  glsl::uimage2D hash2;
  float amb[3], dif[3], spe[3];
};

struct RunArgs {
  const float* entry;        // W*H*4: ray entry (normalised volume coords, eye-space z) -- rayStartPoint of a new view
  const float* exit_eye;     // W*H*3: vPosInViewCoords, the bbox back-face position in eye space
  const uint8_t* covered;    // W*H: the back faces cover this pixel (the fragment shader runs)
  float *ray_pos, *ray_aux, *image, *hit_pos, *hit_normal;
  uint32_t x0, y0, x1, y1;
};
