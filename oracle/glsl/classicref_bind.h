// TEST INFRASTRUCTURE ONLY -- what the CPU harness binds to the reference's classic ray-caster shader programs
// (see build_classicref.py).
#pragma once
#include "glsl_runtime.h"

namespace glsl {
struct mat3 {                              // column major, as gl_NormalMatrix
  float m[9];
  mat3() { std::memset(m, 0, sizeof(m)); }
};
inline vec3 operator*(const mat3& M, const vec3& v) {
  vec3 r;
  for (int row = 0; row < 3; ++row) r.d[row] = (M.m[row] * v.d[0] + M.m[3 + row] * v.d[1]) + M.m[6 + row] * v.d[2];
  return r;
}
}  // namespace glsl

struct Bind {
  glsl::sampler2D entry_tex;               // texRayExitPos: the RGBA16F ray-entry buffer (values already rounded to half)
  glsl::sampler1D tf1;
  glsl::sampler2D tf2;
  glsl::sampler3D vol;                     // the brick's own 3-D texture
  glsl::sampler2D hit_pos, hit_normal, hit_pos2, hit_normal2;
  float fTransScale = 0, fStepScale = 0, fRayStepsize = 0, fGradientScale = 0, fIsoval = 0;
  int iTileID = 0;
  glsl::vec2 vScreensize, vProjParam;
  glsl::vec3 vVoxelStepsize, vDomainScale, vLightAmbient, vLightDiffuse, vLightDiffuse2, vLightSpecular, vLightDir, vCVParam, vCVPickPos;
  glsl::mat4 texmat;
  glsl::mat3 normalmat;
  glsl::vec3 eyepos;
  glsl::vec4 fragcoord;
  glsl::vec4 texcoord;                     // gl_TexCoord[0]
  glsl::sampler2D last, left_eye, right_eye;
  float fSplitCoord = 0.5f;
  int iAlternatingFrameID = 0;
};

struct Out {
  glsl::vec4 color, data0, data1;
  float depth = 0;
};
