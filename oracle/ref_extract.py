"""TEST INFRASTRUCTURE ONLY -- runs STATEMENTS OF THE REFERENCE that sit inside functions too entangled with OpenGL / Qt to
compile whole: the statements are cut out of the .cpp where it lies under /root/reference (brace / marker matching, as
oracle/glsl/build_glslref.py does for the shader generators), wrapped in a stub that supplies only the names they read,
and compiled into oracle/_ref/ref_extract.  Nothing of the reference is copied into the repository.

  poolsize      GPUMemMan::GetVolumePool's pool sizing, Renderer/GPUMemMan/GPUMemMan.cpp:769-830
  histresample  the 1-D histogram "resampling" loop of UVFDataset::GetHistograms, IO/uvfDataset.cpp:762-798
  maxlod        AbstrRenderer::ComputeMaxLODForCurrentView (whole body), Renderer/AbstrRenderer.cpp:681-787
  slice2d       the set-up of a 2-D slice view in GLRenderer::Render2DView (texture window of the single brick, aspect ratios,
                slice coordinate), Renderer/GL/GLRenderer.cpp:1129-1171
  hqmiplod      the level-of-detail rule of AbstrRenderer::PlanHQMIPFrame, Renderer/AbstrRenderer.cpp:1220-1235
  mipmatrices   (second program, oracle/_ref/ref_extract_gl: Vectors.h's USEGL section is needed for MATRIX4::Ortho) the
                orthographic projection of the MIP view, Renderer/GL/GLRenderer.cpp:1184-1196, and the whole body of
                GLRenderer::RenderHQMIPPreLoop, :1256-1283 (m_maMIPRotation); and the whole body of GLRenderer::RenderSlice,
                :922-1014, with glTexCoord3d / glVertex3d recording the quad it draws
"""
import os
import subprocess
import sys
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "glsl"))
import build_glslref as B          # noqa: E402  (cut_body / read and the include flags)

OUT = os.path.join(_HERE, "_ref")
BIN = os.path.join(OUT, "ref_extract")


def available():
    return os.path.isfile(os.path.join(B.T, "Renderer/GPUMemMan/GPUMemMan.cpp"))


def _pool_statements():
    mm = B.read(os.path.join(B.T, "Renderer/GPUMemMan/GPUMemMan.cpp"))
    body = B.cut_body(mm, r"GLVolumePool\* GPUMemMan::GetVolumePool\(")
    return body[body.index("const uint64_t iBitWidth"):body.index("GLVolumePool* pPool = NULL;")]


def _hist_statements():
    ds = B.read(os.path.join(B.T, "IO/uvfDataset.cpp"))
    body = B.cut_body(ds, r"void UVFDataset::GetHistograms\(size_t\)")
    a = body.index("float sampleFactor = ")
    copy_loop = body.index("for (size_t i = 0;i < m_pHist1D->GetSize(); i++)", a)
    return body[a:body.rindex("} else {", a, copy_loop)]


def _maxlod_body():
    ar = B.read(os.path.join(B.T, "Renderer/AbstrRenderer.cpp"))
    return B.cut_body(ar, r"void AbstrRenderer::ComputeMaxLODForCurrentView\(\)")


def _slice2d_statements():
    gl = B.read(os.path.join(B.T, "Renderer/GL/GLRenderer.cpp"))
    body = B.cut_body(gl, r"bool GLRenderer::Render2DView\(RenderRegion2D& renderRegion\)")
    a = body.index("UINTVECTOR3 vRealVoxelCount;")
    b = body.index("const int sliceDir = static_cast<size_t>(renderRegion.windowMode);")
    c = body.index("double fSliceIndex = static_cast<double>(renderRegion.GetSliceIndex())")
    d = body.index("RenderSlice(renderRegion, fSliceIndex", c)
    return body[a:b], body[c:d]


def _hqmiplod_statements():
    ar = B.read(os.path.join(B.T, "Renderer/AbstrRenderer.cpp"))
    body = B.cut_body(ar, r"void AbstrRenderer::PlanHQMIPFrame\(RenderRegion& renderRegion\)")
    return body[body.index("UINTVECTOR3  viVoxelCount"):body.index("// build new brick todo-list")]


def _f3_source():
    first, second = _slice2d_statements()
    return (r"""// ref_extract slice2d <domain xyz> <lod domain xyz> <brick voxels xyz> <scale xyz> <win w h> <window mode> <slice index>
namespace MathTools { static inline uint32_t NextPow2(uint32_t v) { uint32_t p = 1; while (p < v) p <<= 1; return p; } }   // (branch not taken: m_bUseOnlyPowerOfTwo is false)
struct Stub2DDataset {
  UINT64VECTOR3 dom, lod; DOUBLEVECTOR3 scale; uint64_t lodCount;
  UINT64VECTOR3 GetDomainSize(size_t l = 0) const { return l ? lod : dom; } DOUBLEVECTOR3 GetScale() const { return scale; }
  uint64_t GetLODLevelCount() const { return lodCount; }
};
struct StubRegion2D { UINTVECTOR2 minCoord, maxCoord; int windowMode; uint64_t slice; uint64_t GetSliceIndex() const { return slice; } };
static int slice2d(char** a) {
  Stub2DDataset ds; Stub2DDataset* m_pDataset = &ds;
  ds.dom = UINT64VECTOR3(strtoull(a[2], 0, 10), strtoull(a[3], 0, 10), strtoull(a[4], 0, 10));
  ds.lod = UINT64VECTOR3(strtoull(a[5], 0, 10), strtoull(a[6], 0, 10), strtoull(a[7], 0, 10));
  UINTVECTOR3 vVoxelCount((uint32_t)strtoul(a[8], 0, 10), (uint32_t)strtoul(a[9], 0, 10), (uint32_t)strtoul(a[10], 0, 10));
  ds.scale = DOUBLEVECTOR3(atof(a[11]), atof(a[12]), atof(a[13]));
  StubRegion2D renderRegion; renderRegion.minCoord = UINTVECTOR2(0, 0);
  renderRegion.maxCoord = UINTVECTOR2((uint32_t)strtoul(a[14], 0, 10), (uint32_t)strtoul(a[15], 0, 10));
  renderRegion.windowMode = atoi(a[16]); renderRegion.slice = strtoull(a[17], 0, 10);
  const bool m_bUseOnlyPowerOfTwo = false; const size_t iCurrentLOD = 1;
""" + first + """
  const int sliceDir = renderRegion.windowMode;
""" + second + r"""
  printf("%.9g %.9g %.9g %.9g %.9g %.9g %.17g %.17g %.17g %.17g %.17g %.17g\n", vMinCoords.x, vMinCoords.y, vMinCoords.z, vMaxCoords.x, vMaxCoords.y,
         vMaxCoords.z, vAspectRatio.x, vAspectRatio.y, vAspectRatio.z, vWinAspectRatio.x, vWinAspectRatio.y, fSliceIndex);
  return 0;
}
// ref_extract q <domain xyz> <win w h> <lod count> <mip lod 0|1>
static int hqmiplod(char** a) {
  Stub2DDataset ds; Stub2DDataset* m_pDataset = &ds;
  ds.dom = UINT64VECTOR3(strtoull(a[2], 0, 10), strtoull(a[3], 0, 10), strtoull(a[4], 0, 10));
  const UINTVECTOR2 m_vWinSize((uint32_t)strtoul(a[5], 0, 10), (uint32_t)strtoul(a[6], 0, 10));
  ds.lodCount = strtoull(a[7], 0, 10);
  const bool m_bMIPLOD = atoi(a[8]) != 0;
  uint64_t m_iCurrentLODOffset = 0, m_iCurrentLOD = 0;
  using std::min;
""" + _hqmiplod_statements() + r"""
  printf("%llu\n", (unsigned long long)m_iCurrentLOD);
  return 0;
}
""")


def _source():
    return (r'''// GENERATED by oracle/ref_extract.py from the reference's GPUMemMan.cpp / uvfDataset.cpp -- do not commit
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "Basics/Vectors.h"
typedef int GLint;
static uint64_t g_gpumem; static int g_max3d;
struct StubSysInfo { uint64_t GetMaxUsableGPUMem() const { return g_gpumem; } };
struct StubController { static StubController& ConstInstance() { static StubController c; return c; } StubSysInfo SysInfo() const { return StubSysInfo(); } };
struct StubDataset {
  uint64_t bits, comps, total; UINT64VECTOR3 bs;
  uint64_t GetBitWidth() const { return bits; } uint64_t GetComponentCount() const { return comps; }
  UINT64VECTOR3 GetMaxUsedBrickSizes() const { return bs; } uint64_t GetTotalBrickCount() const { return total; }
};
struct StubHist1D { std::vector<uint32_t> d; void Set(size_t i, uint32_t v) { d[i] = v; } size_t GetSize() const { return d.size(); } };
#define MESSAGE(...)

// ref_extract hist <in.u64> <n> <target> <out.u32>
static int hist_resample(char** argv) {
  const size_t n = strtoull(argv[3], 0, 10), target = strtoull(argv[4], 0, 10);
  std::vector<uint64_t> vHist1D(n);
  FILE* f = fopen(argv[2], "rb");
  if (!f || fread(vHist1D.data(), 8, n, f) != n) return 1;
  fclose(f);
  StubHist1D h; h.d.assign(target, 0u);                    // (the reference's new[] leaves the never-written last bin undefined)
  StubHist1D* m_pHist1D = &h;
  {
''' + _hist_statements() + r'''
  }
  f = fopen(argv[5], "wb");
  fwrite(h.d.data(), 4, target, f);
  fclose(f);
  return 0;
}

// ref_extract maxlod: reads state lines "target ms0 ms1 maxms notok skip lowres decrate decres listsize rendered maxidx minlod limx"
// from stdin, runs the reference body on each and prints the state back followed by "start current"
enum ERendererTarget { RT_INTERACTIVE = 0, RT_CAPTURE, RT_HEADLESS };
struct StubRenderer {
  ERendererTarget m_eRendererTarget; float msecPassed[2]; float m_fMaxMSPerFrame; uint32_t m_iLODNotOKCounter;
  uint64_t m_iPerformanceBasedLODSkip; bool m_bRenderLowResIntermediateResults, decreaseSamplingRate, decreaseScreenRes;
  std::vector<int> m_vCurrentBrickList; uint64_t m_iBricksRenderedInThisSubFrame, m_iMaxLODIndex, m_iMinLODForCurrentView;
  uint64_t m_iStartLODOffset, m_iCurrentLODOffset; UINTVECTOR2 m_iLODLimits;
  void RestartTimers() { msecPassed[0] = msecPassed[1] = -1.0f; }
  void ComputeMaxLODForCurrentView();
};
void StubRenderer::ComputeMaxLODForCurrentView() {
@@MAXLOD@@
}
static int maxlod_loop() {
  StubRenderer r;
  int target, lowres, decrate, decres; unsigned long long skip, listsize, rendered, maxidx, minlod, limx; unsigned notok;
  while (scanf("%d %f %f %f %u %llu %d %d %d %llu %llu %llu %llu %llu", &target, &r.msecPassed[0], &r.msecPassed[1], &r.m_fMaxMSPerFrame, &notok,
               &skip, &lowres, &decrate, &decres, &listsize, &rendered, &maxidx, &minlod, &limx) == 14) {
    r.m_eRendererTarget = (ERendererTarget)target; r.m_iLODNotOKCounter = notok; r.m_iPerformanceBasedLODSkip = skip;
    r.m_bRenderLowResIntermediateResults = lowres != 0; r.decreaseSamplingRate = decrate != 0; r.decreaseScreenRes = decres != 0;
    r.m_vCurrentBrickList.assign((size_t)listsize, 0); r.m_iBricksRenderedInThisSubFrame = rendered; r.m_iMaxLODIndex = maxidx;
    r.m_iMinLODForCurrentView = minlod; r.m_iLODLimits = UINTVECTOR2((uint32_t)limx, 0); r.m_iStartLODOffset = 0; r.m_iCurrentLODOffset = 0;
    r.ComputeMaxLODForCurrentView();
    printf("%.9g %.9g %u %llu %d %d %llu %llu\n", r.msecPassed[0], r.msecPassed[1], r.m_iLODNotOKCounter,
           (unsigned long long)r.m_iPerformanceBasedLODSkip, (int)r.decreaseSamplingRate, (int)r.decreaseScreenRes,
           (unsigned long long)r.m_iStartLODOffset, (unsigned long long)r.m_iCurrentLODOffset);
  }
  return 0;
}

@@F3@@
#define Controller StubController
#define m_iAllocatedGPUMemory uint64_t(0)
#define GL(stmt) iMaxVolumeDims = g_max3d
// ref_extract <bit width> <components> <bsx> <bsy> <bsz> <total bricks> <max 3-D texture> <gpu bytes>
int main(int argc, char** argv) {
  if (argc == 6 && argv[1][0] == 'h') return hist_resample(argv);
  if (argc == 2 && argv[1][0] == 'm') return maxlod_loop();
  if (argc == 18 && argv[1][0] == 's') return slice2d(argv);
  if (argc == 9 && argv[1][0] == 'q') return hqmiplod(argv);
  if (argc != 9) return 2;
  StubDataset ds; StubDataset* dataSet = &ds;
  ds.bits = strtoull(argv[1], 0, 10); ds.comps = strtoull(argv[2], 0, 10);
  ds.bs = UINT64VECTOR3(strtoull(argv[3], 0, 10), strtoull(argv[4], 0, 10), strtoull(argv[5], 0, 10));
  ds.total = strtoull(argv[6], 0, 10); g_max3d = atoi(argv[7]); g_gpumem = strtoull(argv[8], 0, 10);
''' + _pool_statements() + r'''
  printf("%u %u %u\n", poolSize.x, poolSize.y, poolSize.z);
  return 0;
}
''').replace("@@MAXLOD@@", _maxlod_body()).replace("@@F3@@", _f3_source())


def build():
    os.makedirs(OUT, exist_ok=True)
    src = _source()
    cpp = os.path.join(OUT, "ref_extract.cpp")
    if not os.path.exists(BIN) or not os.path.exists(cpp) or B.read(cpp) != src:
        with open(cpp, "w") as f:
            f.write(src)
        shim = os.path.join(_HERE, "ref_build", "shim")
        subprocess.check_call(["g++", "-O1", "-std=c++14", "-fpermissive", "-w", "-DTUVOK_NO_QT=1", "-I" + shim, "-I" + B.T,
                               "-I" + os.path.join(B.T, "Basics"), "-I" + os.path.join(B.T, "IO/3rdParty/boost"), cpp, "-o", BIN])
    return BIN


def pool_size(bit_width, max_brick, total_bricks, max_3d_texture, gpu_mem_bytes, components=1):
    """poolSize (voxels per axis) as the reference's GetVolumePool computes it"""
    out = subprocess.check_output([build(), str(bit_width), str(components), str(max_brick[0]), str(max_brick[1]), str(max_brick[2]),
                                   str(total_bricks), str(max_3d_texture), str(gpu_mem_bytes)]).decode().split()
    return tuple(int(v) for v in out)


def hist_resample(bins, target):
    """the reference's resampling loop on `bins` (uint64) -> `target` uint32 bins"""
    import numpy as np
    with tempfile.TemporaryDirectory() as d:
        a, b = os.path.join(d, "in.u64"), os.path.join(d, "out.u32")
        np.ascontiguousarray(bins, np.uint64).tofile(a)
        subprocess.check_call([build(), "hist", a, str(len(bins)), str(int(target)), b])
        return np.fromfile(b, np.uint32)


def max_lod_for_view(states):
    """states: list of 14-tuples (target, ms0, ms1, max_ms, not_ok, skip, low_res, dec_rate, dec_res, list_size, rendered, max_lod_index,
    min_lod, lod_limit_x) -> list of (ms0, ms1, not_ok, skip, dec_rate, dec_res, start, current) after the reference's function"""
    text = "\n".join(" ".join(repr(float(v)) if isinstance(v, float) else str(int(v)) for v in st) for st in states) + "\n"
    out = subprocess.run([build(), "maxlod"], input=text.encode(), stdout=subprocess.PIPE, check=True).stdout.decode().split("\n")
    res = []
    for line in out:
        if line.strip():
            f = line.split()
            res.append((float(f[0]), float(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6]), int(f[7])))
    return res


def slice2d(domain, lod_domain, brick_voxels, scale, win, window_mode, slice_index):
    """the reference's own statements of GLRenderer::Render2DView for a slice view -> dict(min_coords, max_coords, aspect, win_aspect, slice)"""
    args = [str(int(v)) for v in domain] + [str(int(v)) for v in lod_domain] + [str(int(v)) for v in brick_voxels] + \
           [repr(float(v)) for v in scale] + [str(int(win[0])), str(int(win[1])), str(int(window_mode)), str(int(slice_index))]
    f = [float(t) for t in subprocess.check_output([build(), "slice2d"] + args).decode().split()]
    return dict(min_coords=f[0:3], max_coords=f[3:6], aspect=f[6:9], win_aspect=f[9:11], slice=f[11])


def hq_mip_lod(domain, win, lod_count, mip_lod=True):
    """m_iCurrentLOD as AbstrRenderer::PlanHQMIPFrame's statements leave it"""
    args = [str(int(v)) for v in domain] + [str(int(win[0])), str(int(win[1])), str(int(lod_count)), "1" if mip_lod else "0"]
    return int(subprocess.check_output([build(), "q"] + args).decode().split()[0])


BIN_GL = os.path.join(OUT, "ref_extract_gl")


def _gl_source():
    gl = B.read(os.path.join(B.T, "Renderer/GL/GLRenderer.cpp"))
    body = B.cut_body(gl, r"bool GLRenderer::Render2DView\(RenderRegion2D& renderRegion\)")
    a = body.index("FLOATMATRIX4 maOrtho;")
    ortho = body[a:body.index("maOrtho.setProjection();", a)]
    preloop = B.cut_body(gl, r"void GLRenderer::RenderHQMIPPreLoop\(RenderRegion2D& region\)")
    return (r"""// GENERATED by oracle/ref_extract.py from the reference's GLRenderer.cpp -- do not commit
#include <algorithm>
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
#define T_ERROR(...)
using std::max;
struct RenderRegion { enum EWindowMode { WM_SAGITTAL = 0, WM_AXIAL = 1, WM_CORONAL = 2, WM_3D, WM_INVALID }; };
struct RenderRegion2D { int windowMode; VECTOR2<bool> flipView; FLOATMATRIX4 rotation; };
struct StubGLRenderer {
  float m_fMIPRotationAngle; FLOATMATRIX4 m_maMIPRotation;
  void RenderHQMIPPreLoop(RenderRegion2D& region);
  void RenderSlice(const RenderRegion2D& region, double fSliceIndex, FLOATVECTOR3 vMinCoords, FLOATVECTOR3 vMaxCoords,
                   DOUBLEVECTOR3 vAspectRatio, DOUBLEVECTOR2 vWinAspectRatio);
};
enum { GL_QUADS = 7 };
static double g_tc[3]; static int g_nv = 0;
static void glBegin(int) {}
static void glEnd() {}
static void glTexCoord3d(double a, double b, double c) { g_tc[0] = a; g_tc[1] = b; g_tc[2] = c; }
static void glVertex3d(double x, double y, double z) { printf("%.9g %.9g %.9g %.9g %.9g\n", x, y, g_tc[0], g_tc[1], g_tc[2]); ++g_nv; }
void StubGLRenderer::RenderSlice(const RenderRegion2D& region, double fSliceIndex, FLOATVECTOR3 vMinCoords, FLOATVECTOR3 vMaxCoords,
                                 DOUBLEVECTOR3 vAspectRatio, DOUBLEVECTOR2 vWinAspectRatio) {
""" + B.cut_body(gl, r"void GLRenderer::RenderSlice\(const RenderRegion2D& region, double fSliceIndex,") + r"""
}
void StubGLRenderer::RenderHQMIPPreLoop(RenderRegion2D& region) {
""" + preloop + r"""
}
// ref_extract_gl <win w> <win h> <window mode> <flip x> <flip y> <angle>: the orthographic projection, then m_maMIPRotation
int main(int argc, char** argv) {
  if (argc == 17) {   // slice <mode> <flip x> <flip y> <slice> <min xyz> <max xyz> <aspect xyz> <win aspect xy>: the quad RenderSlice draws
    RenderRegion2D region; region.windowMode = atoi(argv[2]); region.flipView = VECTOR2<bool>(atoi(argv[3]) != 0, atoi(argv[4]) != 0);
    StubGLRenderer r;
    r.RenderSlice(region, atof(argv[5]), FLOATVECTOR3((float)atof(argv[6]), (float)atof(argv[7]), (float)atof(argv[8])),
                  FLOATVECTOR3((float)atof(argv[9]), (float)atof(argv[10]), (float)atof(argv[11])),
                  DOUBLEVECTOR3(atof(argv[12]), atof(argv[13]), atof(argv[14])), DOUBLEVECTOR2(atof(argv[15]), atof(argv[16])));
    return g_nv == 4 ? 0 : 1;
  }
  if (argc != 7) return 2;
  const UINTVECTOR2 m_vWinSize((unsigned)atoi(argv[1]), (unsigned)atoi(argv[2]));
""" + ortho + r"""
  StubGLRenderer r; r.m_fMIPRotationAngle = (float)atof(argv[6]);
  RenderRegion2D region; region.windowMode = atoi(argv[3]); region.flipView = VECTOR2<bool>(atoi(argv[4]) != 0, atoi(argv[5]) != 0);
  r.RenderHQMIPPreLoop(region);
  for (int i = 0; i < 16; ++i) printf("%.9g ", maOrtho.array[i]);
  printf("\n");
  for (int i = 0; i < 16; ++i) printf("%.9g ", r.m_maMIPRotation.array[i]);
  printf("\n");
  return 0;
}
""")


def build_gl():
    os.makedirs(OUT, exist_ok=True)
    src = _gl_source()
    cpp = os.path.join(OUT, "ref_extract_gl.cpp")
    if not os.path.exists(BIN_GL) or not os.path.exists(cpp) or B.read(cpp) != src:
        with open(cpp, "w") as f:
            f.write(src)
        subprocess.check_call(["g++", "-O1", "-std=c++14", "-fpermissive", "-w", "-I" + B.T, "-I" + os.path.join(B.T, "Basics"), cpp, "-o", BIN_GL])
    return BIN_GL


def mip_matrices(win, window_mode, flip, angle_deg):
    """(orthographic projection, m_maMIPRotation) as the reference's statements compute them, 4x4 float32 each (row-vector storage)"""
    import numpy as np
    out = subprocess.check_output([build_gl(), str(int(win[0])), str(int(win[1])), str(int(window_mode)), "1" if flip[0] else "0",
                                   "1" if flip[1] else "0", repr(float(angle_deg))]).decode().strip().splitlines()
    rows = [np.array([float(t) for t in line.split()], np.float32).reshape(4, 4) for line in out]
    return rows[0], rows[1]


def slice_quad(window_mode, flip, slice_coord, min_coords, max_coords, aspect, win_aspect):
    """the quad GLRenderer::RenderSlice draws, from the reference's own text: four (x, y, s, t, r) rows -- vertex position in
    normalised device coordinates and its 3-D texture coordinate"""
    import numpy as np
    args = ["slice", str(int(window_mode)), "1" if flip[0] else "0", "1" if flip[1] else "0", repr(float(slice_coord))] + \
           [repr(float(v)) for v in min_coords] + [repr(float(v)) for v in max_coords] + [repr(float(v)) for v in aspect] + \
           [repr(float(v)) for v in win_aspect]
    out = subprocess.check_output([build_gl()] + args).decode().strip().splitlines()
    return np.array([[float(t) for t in line.split()] for line in out], np.float64)
