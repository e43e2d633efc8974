// TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// stereo_ref <eye xyz> <at xyz> <up xyz> <fovy> <aspect> <znear> <zfar> <focalLength> <eyeDist>
//   calls the reference's FLOATMATRIX4::BuildStereoLookAtAndProjection (Basics/Vectors.h:1215-1248) the way
//   GLRenderer::ComputeViewAndProjection does (GLRenderer.cpp:905-909) and prints view left, view right, projection left,
//   projection right (16 floats each, in the reference's storage order).  The function sits inside Vectors.h's USEGL
//   section, so this file declares the handful of fixed-function GL entry points that section mentions (never called
//   here) and includes the header unmodified from /root/reference.  Built by oracle/ref_build/Makefile into oracle/_ref/.
#include <cstdio>
#include <cstdlib>
typedef float GLfloat;
typedef unsigned GLenum;
enum { GL_MODELVIEW = 1, GL_MODELVIEW_MATRIX, GL_PROJECTION, GL_PROJECTION_MATRIX, GL_TEXTURE, GL_TEXTURE0_ARB };
static void glGetFloatv(GLenum, GLfloat*) { abort(); }
static void glLoadMatrixf(const GLfloat*) { abort(); }
static void glMultMatrixf(const GLfloat*) { abort(); }
static void glMatrixMode(GLenum) { abort(); }
static void glActiveTextureARB(GLenum) { abort(); }
static void glNormal3f(GLfloat, GLfloat, GLfloat) { abort(); }
static void glVertex2f(GLfloat, GLfloat) { abort(); }
static void glVertex3f(GLfloat, GLfloat, GLfloat) { abort(); }
static void glVertex4f(GLfloat, GLfloat, GLfloat, GLfloat) { abort(); }
#define USEGL
#include "Basics/Vectors.h"

int main(int argc, char** argv) {
  if (argc < 16) { fprintf(stderr, "usage: stereo_ref eye(3) at(3) up(3) fovy aspect znear zfar focalLength eyeDist\n"); return 2; }
  float a[15];
  for (int i = 0; i < 15; ++i) a[i] = strtof(argv[1 + i], nullptr);
  FLOATMATRIX4 v[2], p[2];
  FLOATMATRIX4::BuildStereoLookAtAndProjection(FLOATVECTOR3(a[0], a[1], a[2]), FLOATVECTOR3(a[3], a[4], a[5]), FLOATVECTOR3(a[6], a[7], a[8]),
                                               a[9], a[10], a[11], a[12], a[13], a[14], v[0], v[1], p[0], p[1]);
  const FLOATMATRIX4* all[4] = {&v[0], &v[1], &p[0], &p[1]};
  for (int m = 0; m < 4; ++m) {
    for (int i = 0; i < 16; ++i) printf("%.9g ", all[m]->array[i]);
    printf("\n");
  }
  return 0;
}
