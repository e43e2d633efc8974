"""TEST INFRASTRUCTURE ONLY -- builds oracle/_ref/glslref/classic/libclassicref.so: the reference's OWN shaders of the classic
per-brick ray-caster (GLRaycaster) compiled for the CPU, the way build_glslref.py does it for the ray-guided renderer.

What runs inside is the reference's shader text, read at build time from /root/reference/Tuvok/Shaders, linked per program
as GLRaycaster::LoadShaders lists the objects (Renderer/GL/GLRaycaster.cpp:108-192):
  1D      Compositing + Volume3D + VRender1D + VRender1D-BScale + VRender1DProxy + GLRaycaster-1D-FS
  1D lit  Compositing + Volume3D + lighting + VRender1DLit + GLRaycaster-1D-light-FS
  2D      Compositing + Volume3D + GLRaycaster-2D-FS            2D lit: + lighting, GLRaycaster-2D-light-FS
  iso     RefineIsosurface + Volume3D + GLRaycaster-ISO-FS      ClearView: the same with GLRaycaster-ISO-CV-FS
  MIP     Volume3D + GLRaycaster-MIP-Rot-FS
  compose Compose-FS / Compose-CV-FS (GLRenderer.cpp:317-336)
These are compatibility-profile shaders: besides the storage-syntax rewrites of build_glslref.to_cpp, `varying` inputs,
gl_TextureMatrix, gl_NormalMatrix and gl_FragData become globals and texture1D / texture3D are spelled texture.
Around the shader text, restated (oracle/glsl/classicref_entry.inc): which pixels a brick's faces cover and the
interpolated positions (analytic ray / box intersection, entry rounded to half precision as the RGBA16F buffer does), the
uniform values, the frame-buffer blend between bricks and the depth tests.  Outputs go to oracle/_ref/ only."""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import build_glslref as B   # noqa: E402

OUT = os.path.join(B.OUT, "classic")

PROGRAMS = {
    "c1d": ["Compositing.glsl", "Volume3D.glsl", "VRender1D.glsl", "VRender1D-BScale.glsl", "VRender1DProxy.glsl", "GLRaycaster-1D-FS.glsl"],
    "c1d_l": ["Compositing.glsl", "Volume3D.glsl", "lighting.glsl", "VRender1DLit.glsl", "GLRaycaster-1D-light-FS.glsl"],
    "c2d": ["Compositing.glsl", "Volume3D.glsl", "GLRaycaster-2D-FS.glsl"],
    "c2d_l": ["Compositing.glsl", "Volume3D.glsl", "lighting.glsl", "GLRaycaster-2D-light-FS.glsl"],
    "ciso": ["RefineIsosurface.glsl", "Volume3D.glsl", "GLRaycaster-ISO-FS.glsl"],
    "ciso_cv": ["RefineIsosurface.glsl", "Volume3D.glsl", "GLRaycaster-ISO-CV-FS.glsl"],
    "cmip": ["Volume3D.glsl", "GLRaycaster-MIP-Rot-FS.glsl"],
    "ccompose": ["Compose-FS.glsl"],
    "ccompose_cv": ["Compose-CV-FS.glsl"],
    # GLRenderer::LoadShaders (GLRenderer.cpp:295-352): the 2-D slice views, the MIP transfer pass, the stereo compositions
    "s1d": ["Volume3D.glsl", "1D-slice-FS.glsl"],
    "s2d": ["Volume3D.glsl", "2D-slice-FS.glsl"],
    "smip": ["Volume3D.glsl", "MIP-slice-FS.glsl"],
    "tmip": ["Transfer-MIP-FS.glsl"],
    "st_rb": ["Compose-Anaglyphs-FS.glsl"],
    "st_scan": ["Compose-Scanline-FS.glsl"],
    "st_sbs": ["Compose-SBS-FS.glsl"],
    "st_af": ["Compose-AF-FS.glsl"],
}

UNIFORMS = {      # uniform name -> expression over the harness's Bind `b`
    "texRayExitPos": "b.entry_tex", "texRayExit": "b.entry_tex", "texVolume": "b.vol", "texTrans": None,
    "texLastHit": "b.hit_pos", "texLastHitPos": "b.hit_normal", "texRayHitPos": "b.hit_pos", "texRayHitNormal": "b.hit_normal",
    "texRayHitPos2": "b.hit_pos2", "texRayHitNormal2": "b.hit_normal2",
    "fTransScale": "b.fTransScale", "fStepScale": "b.fStepScale", "fRayStepsize": "b.fRayStepsize", "fGradientScale": "b.fGradientScale",
    "fIsoval": "b.fIsoval", "iTileID": "b.iTileID", "ScaleMethod": "0", "TFuncBias": "0.0f",
    "vScreensize": "b.vScreensize", "vProjParam": "b.vProjParam", "vVoxelStepsize": "b.vVoxelStepsize", "vDomainScale": "b.vDomainScale",
    "vLightAmbient": "b.vLightAmbient", "vLightDiffuse": "b.vLightDiffuse", "vLightDiffuse2": "b.vLightDiffuse2",
    "vLightSpecular": "b.vLightSpecular", "vLightDir": "b.vLightDir", "vCVParam": "b.vCVParam", "vCVPickPos": "b.vCVPickPos",
    "texLast": "b.last", "texLeftEye": "b.left_eye", "texRightEye": "b.right_eye", "fSplitCoord": "b.fSplitCoord",
    "iAlternatingFrameID": "b.iAlternatingFrameID",
}


def old_style(src):
    """compatibility-profile spellings -> what build_glslref.to_cpp and glsl_runtime.h know"""
    src = B.strip_comments(src)
    src = re.sub(r"\bvarying\s+vec3\s+vEyePos\s*;", "", src)
    src = re.sub(r"\btexture[13]D\s*\(", "texture(", src)
    # a swizzle passed to an `inout` parameter: GLSL copies in and out, a C++ reference cannot bind to it -- spell the copies
    src = re.sub(r"(\w+)\.xyz\s*=\s*RefineIsosurface\(\s*(\w+)\s*,\s*\1\.xyz\s*,\s*(\w+)\s*\)\s*;",
                 r"{ vec3 io_arg = \1.xyz; vec3 io_ret = RefineIsosurface(\2, io_arg, \3); \1.xyz = io_arg; \1.xyz = io_ret; }", src)
    return src


def program_source(name, objects):
    uniforms, inouts = {}, {}
    body = []
    for o in objects:
        body.append("// ---- %s\n" % o + B.to_cpp(old_style(B.read(os.path.join(B.T, "Shaders", o))), uniforms, inouts))
    decl = ["static %s %s;" % (t, n) for n, t in sorted(uniforms.items())]
    decl += ["static vec4 gl_FragCoord;", "static vec4 gl_FragColor;", "static float gl_FragDepth;", "static vec4 gl_FragData[4];",
             "static mat4 gl_TextureMatrix[1];", "static mat3 gl_NormalMatrix;", "static vec3 vEyePos;", "static vec4 gl_TexCoord[1];"]
    sets = []
    for n, t in sorted(uniforms.items()):
        if n == "texTrans":
            sets.append("  texTrans = b.tf%s;" % ("1" if t == "sampler1D" else "2"))
        elif n in UNIFORMS:
            sets.append("  %s = %s;" % (n, UNIFORMS[n]))
        else:
            raise ValueError("shader program %s declares uniform %s %s that the harness does not know" % (name, t, n))
    return ("namespace prog_%s {\nusing namespace glsl;\n" % name + "\n".join(decl) + "\n" + "\n".join(body) +
            "\nstatic bool run1(const Bind& b, Out& o) {\n" + "\n".join(sets) +
            "\n  vEyePos = b.eyepos; gl_TextureMatrix[0] = b.texmat; gl_NormalMatrix = b.normalmat; gl_FragCoord = b.fragcoord; gl_TexCoord[0] = b.texcoord;\n"
            "  gl_FragColor = vec4(0.0f); gl_FragDepth = 0.0f; for (int k = 0; k < 4; ++k) gl_FragData[k] = vec4(0.0f);\n"
            "  try { shader_main(); } catch (const discard_exception&) { return false; }\n"
            "  o.color = gl_FragColor; o.data0 = gl_FragData[0]; o.data1 = gl_FragData[1]; o.depth = gl_FragDepth;\n  return true;\n}\n}\n")


def build():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libclassicref.so")
    unit = ['// GENERATED by oracle/glsl/build_classicref.py from the reference shaders -- do not commit',
            '#include "glsl_runtime.h"', '#include "tvk_cuda.h"', '#include "classicref_bind.h"']
    for name, objects in PROGRAMS.items():
        unit.append(program_source(name, objects))
    unit.append('#include "classicref_entry.inc"')
    cpp = os.path.join(OUT, "classicref_unit.cpp")
    with open(cpp, "w") as f:
        f.write("\n".join(unit))
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared", "-w",
                           "-I" + HERE, "-I" + os.path.join(B.ROOT, "include"), cpp, "-o", lib])
    return lib


if __name__ == "__main__":
    print(build())
