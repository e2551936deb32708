"""The constants and GL state this repository assumes, checked mechanically against the reference's SOURCE TEXT.

The shaders are executed by test_reference_glsl.py; everything else the result depends on is host-side state the
reference's C++ sets (matrices, sizes, texture parameters, cull / depth state).  That C++ cannot be built here, but it
can be read: each assertion below extracts the value from the cited file and compares it with what
`vct_b200.uniforms.reference_uniforms()` (the library's defaults), the glue in tests/glsl_harness.py and the drop-in
facade `host/Voxel_Cone_Tracing.h` use.  Skipped where /root/reference is absent (the GPU box).
"""
import os
import re

import numpy as np
import pytest

import glsl_harness as gh
import glsl_run
from vct_b200 import glmath as gm
from vct_b200 import uniforms

REF = "/root/reference/Voxel_Cone_Tracing_Final"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="the reference sources are not on this machine")
NUM = r"(-?\d+(?:\.\d*)?)f?"


def text(name):
    return open(os.path.join(REF, name), errors="ignore").read()


def floats(pattern, src, n):
    m = re.search(pattern, src)
    assert m, pattern
    return [float(m.group(k + 1)) for k in range(n)]


def section(src, start, end):
    a = src.index(start)
    return src[a:src.index(end, a)]


# -------------------------------------------------------------------------------------- shader constants
def literal_args(prog, func, callee):
    """float / int literals passed to `callee` anywhere in the body of `func` (walks the AST)"""
    found = []

    def walk(n):
        if isinstance(n, tuple):
            if n and n[0] == "call" and n[1] == ("name", callee):
                found.extend(a[1] for a in n[2] if a[0] in ("float", "int"))
            for x in n:
                walk(x)
        elif isinstance(n, list):
            for x in n:
                walk(x)
    walk(prog.funcs[func][4])
    return found


def test_shader_constants_are_the_library_defaults():
    """MAX_DISTANCE / MAX_ALPHA / the cone table (VoxelConeTracing.fs:42-57), the two cone apertures (:191, :216), the
    PCF radius, bias and gains (VoxelConeTracing.fs:132-163, Voxelization.fs:18-52) are uniforms of the library; their
    defaults must be the numbers in the shader text."""
    fs = gh.load("VoxelConeTracing.fs", np.float64)
    u = uniforms.reference_uniforms()
    assert float(fs.globals["MAX_DISTANCE"]) == u["MaxDistance"] and float(fs.globals["MAX_ALPHA"]) == u["MaxAlpha"]
    assert fs.globals["NUM_CONES"] == 6 == len(u["ConeWeights"])
    np.testing.assert_array_equal(np.asarray(fs.globals["Cone_Weights"], dtype=np.float32), np.asarray(u["ConeWeights"], dtype=np.float32))
    np.testing.assert_array_equal(np.asarray(fs.globals["Cone_Directions"], dtype=np.float32).reshape(-1),
                                  np.asarray(u["ConeDirections"], dtype=np.float32).reshape(-1))
    assert literal_args(fs, "main", "Voxel_Cone_Tracing") == [u["DiffuseTanHalfAngle"], u["SpecularTanHalfAngle"]]
    assert literal_args(fs, "main", "PCF_Shadow_Mapping") == [u["ShadowBias"]]
    vfs = gh.load("Voxelization.fs", np.float64)
    assert literal_args(vfs, "main", "PCF_Shadow_Mapping") == [u["ShadowBias"]]
    for name in ("VoxelConeTracing.fs", "Voxelization.fs"):
        assert re.search(r"int radius = (\d+);", text("Shader/" + name)).group(1) == str(u["PcfRadius"])
    assert "shadow *= 0.111f;" in text("Shader/VoxelConeTracing.fs")                      # C4: gain 0.111, not 1/25
    assert "shadow /= (2 * radius + 1) * (2 * radius + 1);" in text("Shader/Voxelization.fs")   # V3: true mean
    assert u["ambientFactor"] == floats(r"float AmbientFactor = " + NUM, text("Voxel_Cone_Tracing.h"), 1)[0]


# ---------------------------------------------------------------------------------------- host constants
def test_host_constants_and_matrices_are_the_library_defaults():
    """Voxel_Cone_Tracing.h:14-17, 36, 91-93, 131-137, 164-165, 184-188; Camera.h:21-22, 47; main.cpp:30."""
    h = text("Voxel_Cone_Tracing.h")
    u = uniforms.reference_uniforms()
    assert int(floats(r"const int VoxelDimensions = " + NUM, h, 1)[0]) == u["VoxelDimensions"] == 128
    G = floats(r"const float VoxelGridWorldSize = " + NUM, h, 1)[0]
    assert G == u["VoxelGridWorldSize"] == 150.0
    assert [int(floats(r"int screen_width = " + NUM, h, 1)[0]), int(floats(r"int screen_height = " + NUM, h, 1)[0])] == \
        [u["screen_width"], u["screen_height"]]
    assert int(floats(r"GLuint ShadowMapSize = " + NUM, h, 1)[0]) == u["ShadowMapSize"]
    light = floats(rf"vec3 lightDirection = vec3\({NUM}, {NUM}, {NUM}\)", h, 3)
    assert light == list(u["LightDirection"])
    cam = floats(rf"Camera camera\(glm::vec3\({NUM}, {NUM}, {NUM}\)\)", h, 3)
    assert cam == list(u["CameraPosition"])
    c = text("Camera.h")
    yaw, pitch, zoom = floats(r"const float YAW = " + NUM, c, 1)[0], floats(r"const float PITCH = " + NUM, c, 1)[0], floats(r"float Zoom = " + NUM, c, 1)[0]
    scale = floats(rf"glm::scale\(glm::mat4\(1\.0f\), glm::vec3\({NUM}, {NUM}, {NUM}\)\)", h, 3)
    assert scale == [0.05, 0.05, 0.05]
    # shadow map: lookAt(lightDirection, 0, +y), ortho(-120, 120, -120, 120, -100, 100)           (:91-93)
    o = floats(rf"mat4 pMat = ortho<float>\({NUM}, {NUM}, {NUM}, {NUM}, {NUM}, {NUM}\)", h, 6)
    assert "lookAt(lightDirection, vec3(0.0f, 0.0f, 0.0f), vec3(0.0f, 1.0f, 0.0f))" in h and "DepthViewProjectionMatrix = pMat * vMat;" in h
    model = gm.scale(scale[0])
    depth = gm.ortho(*o) @ gm.look_at(np.asarray(light, dtype=np.float32), (0, 0, 0), (0, 1, 0))
    np.testing.assert_array_equal(gm.colmajor(np.asarray(depth @ model, dtype=np.float32)), u["DepthModelViewProjectionMatrix"])
    # voxelisation: ortho(-size/2, size/2, -size/2, size/2, size/2, 3 size/2) and the three lookAt's      (:131-137)
    assert "ortho<float>(-size * 0.5f, size * 0.5f, -size * 0.5f, size * 0.5f, size * 0.5f, size * 1.5f)" in h
    vp = gm.ortho(-G / 2, G / 2, -G / 2, G / 2, G / 2, G * 1.5)
    for name, eye, up in (("ProjX", "vec3(size, 0.0f, 0.0f)", "vec3(0.0f, 1.0f, 0.0f)"),
                          ("ProjY", "vec3(0.0f, size, 0.0f)", "vec3(0.0f, 0.0f, -1.0f)"),
                          ("ProjZ", "vec3(0.0f, 0.0f, size)", "vec3(0.0f, 0.0f, 0.0f), vec3(0.0f, 1.0f, 0.0f)")):
        assert re.search(re.escape(f"{name} = voxelize_pMat * lookAt({eye}") + ".*" + re.escape(up.split("), ")[-1]), h), name
        e = [G if "size" in part else 0.0 for part in eye[5:-1].split(", ")]
        upv = [float(x.rstrip("f")) for x in up.split("vec3(")[-1].rstrip(")").split(", ")]
        np.testing.assert_array_equal(gm.colmajor(vp @ gm.look_at(e, (0, 0, 0), upv)), u[name])
    # camera: perspective(radians(Zoom), w / h, 0.1, 1000), view from position / YAW / PITCH                (:164-165)
    near, far = floats(rf"\(float\)screen_width / \(float\)screen_height, {NUM}, {NUM}\)", h, 2)
    proj = gm.perspective(gm.radians(zoom), np.float32(u["screen_width"]) / np.float32(u["screen_height"]), near, far)
    np.testing.assert_array_equal(gm.colmajor(proj), u["ProjectionMatrix"])
    view = gm.view_matrix(cam, yaw, pitch)
    np.testing.assert_array_equal(gm.colmajor(np.asarray(view @ model, dtype=np.float32)), u["ModelViewMatrix"])
    assert 'setMat4("ModelViewMatrix", vMat * mMat)' in h and 'setMat4("DepthModelViewProjectionMatrix", DepthViewProjectionMatrix * mMat)' in h
    # the front vector of Camera.h:133-137 is what glmath.view_matrix builds
    assert "front.x = cosf(glm::radians(Yaw)) * cosf(glm::radians(Pitch));" in c and "front.y = sinf(glm::radians(Pitch));" in c
    assert "front.z = sinf(glm::radians(Yaw)) * cosf(glm::radians(Pitch));" in c
    assert re.search(r"glfwWindowHint\(GLFW_SAMPLES, 4\)", text("main.cpp"))              # CoveragePolicy msaa4 default
    assert uniforms.COVERAGE["msaa4"] == u["CoveragePolicy"]


def test_gl_state_the_fixed_function_glue_assumes():
    """tests/glsl_harness.py and the oracle take these from the reference; each is one line of its source."""
    h, m, mesh, main = text("Voxel_Cone_Tracing.h"), text("Model.h"), text("Mesh.h"), text("main.cpp")
    # depth texture: D24, GL_LINEAR both ways, CLAMP_TO_EDGE (shadow_sampler: bilinear, clamp)      (:96-104)
    depth = section(h, "glGenTextures(1, &Depth_Texture);", "glFramebufferTexture")
    assert "GL_DEPTH_COMPONENT24" in depth and depth.count("GL_LINEAR") == 2 and depth.count("GL_CLAMP_TO_EDGE") == 2
    assert "MIPMAP" not in depth
    # voxel texture: RGBA8, LINEAR_MIPMAP_LINEAR / LINEAR, no wrap mode set => GL_REPEAT (Sampler3D wraps)  (:117-131)
    vox = section(h, "glGenTextures(1, &VoxelTexture);", "float size = VoxelGridWorldSize;")
    assert "GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR" in vox and "GL_TEXTURE_MAG_FILTER, GL_LINEAR)" in vox
    assert "GL_TEXTURE_WRAP" not in vox and "GL_RGBA8" in vox and "glGenerateMipmap(GL_TEXTURE_3D);" in vox
    # voxel pass: no culling, no depth test, viewport V x V, image unit RGBA8 write-only, mipmaps after     (:213-251)
    dv = section(h, "void DrawVoxelTexture()", "};")
    assert "glDisable(GL_CULL_FACE);" in dv and "glDisable(GL_DEPTH_TEST);" in dv
    assert "glViewport(0, 0, VoxelDimensions, VoxelDimensions);" in dv and "GL_WRITE_ONLY, GL_RGBA8" in dv
    assert dv.index("model.Draw(VoxelizeShader);") < dv.index("glGenerateMipmap(GL_TEXTURE_3D);")
    # shadow pass and frame: back-face culling, depth test LESS, viewport S x S                             (:192-211)
    dd = section(h, "void DrawDepthTexture()", "void DrawVoxelTexture()")
    assert "glEnable(GL_CULL_FACE);" in dd and "glEnable(GL_DEPTH_TEST);" in dd and "glViewport(0, 0, ShadowMapSize, ShadowMapSize);" in dd
    assert "glDepthFunc(GL_LESS);" in main and "glCullFace(GL_BACK);" in main and "glFrontFace" not in main + h   # CCW front
    rd = section(h, "void Render()", "void DrawDepthTexture()")
    assert "glEnable(GL_CULL_FACE);" in rd and "glEnable(GL_DEPTH_TEST);" in rd
    assert "if (AmbientFactor < 0.5f)\n\t\t\tglClearColor(0.5f, 0.5f, 0.5f, 1.0f);" in rd                          # frame background
    # material textures: REPEAT, trilinear, glGenerateMipmap, channel count -> RED / RGB / RGBA            (Model.h:150-176)
    assert "GL_TEXTURE_WRAP_S, GL_REPEAT" in m and "GL_TEXTURE_WRAP_T, GL_REPEAT" in m
    assert "GL_TEXTURE_MIN_FILTER, GL_LINEAR_MIPMAP_LINEAR" in m and "glGenerateMipmap(GL_TEXTURE_2D);" in m
    assert re.search(r"n == 1\)\s*format = GL_RED;", m) and re.search(r"n == 3\)\s*format = GL_RGB;", m) and re.search(r"n == 4\)\s*format = GL_RGBA;", m)
    assert "stbi_load(file_name.c_str(), &w, &h, &n, 0)" in m and "stbi_set_flip_vertically_on_load" not in m + main
    # per-mesh uniforms: Shininess 20 for everything, HeightTextureSize = the height map's size          (Mesh.h:86-108)
    assert 'glGetUniformLocation(shader.id, "Shininess"), 20.0f' in mesh
    assert '"HeightTextureSize"), textures[i].width, textures[i].height' in mesh
    # vertex layout = the 14 floats vct_upload_mesh takes                                                    (Mesh.h:12-19, 67-79)
    vtx = section(mesh, "struct Vertex", "};")
    assert re.findall(r"vec(\d) (\w+);", vtx) == [("3", "Position"), ("3", "Normal"), ("2", "TexCoords"), ("3", "Tangents"), ("3", "Bi_Tangents")]
    assert "GL_UNSIGNED_INT" in mesh and "GL_TRIANGLES" in mesh


def test_facade_keeps_the_reference_defaults():
    """host/Voxel_Cone_Tracing.h (the drop-in class) declares the same fields with the same initialisers."""
    ref, mine = text("Voxel_Cone_Tracing.h"), open(os.path.join(ROOT, "voxel-cone-tracing_b200", "host", "Voxel_Cone_Tracing.h")).read()
    for pat in (r"lightDirection = vec3\(([^)]*)\)", r"int VoxelDimensions = (\d+)", r"float VoxelGridWorldSize = ([\d.]+f)",
                r"int screen_width = (\d+)", r"int screen_height = (\d+)", r"ShadowMapSize = (\d+)", r"float AmbientFactor = ([\d.]+f)"):
        assert re.search(pat, ref).group(1) == re.search(pat, mine).group(1), pat
    for method in ("init_voxel_cone_tracing", "Render", "DrawDepthTexture", "DrawVoxelTexture"):
        assert re.search(rf"void {method}\(", ref) and re.search(rf"void {method}\(", mine)


def test_interpreter_reads_every_reference_shader():
    """All seven shader files parse; the uniforms / inputs each stage declares are the ones the C ABI names."""
    names = {}
    for f in gh.SHADERS + ["Shadow.fs"]:
        p = glsl_run.Program(text("Shader/" + f))
        names[f] = {n for n, (q, _) in p.decl.items() if "uniform" in q}
    assert names["VoxelConeTracing.vs"] == {"CameraPosition", "ModelMatrix", "ModelViewMatrix", "ProjectionMatrix", "DepthModelViewProjectionMatrix"}
    assert names["Voxelization.gs"] == {"ProjX", "ProjY", "ProjZ"}
    assert {"LightDirection", "ambientFactor", "ShadowMapSize", "VoxelGridWorldSize", "VoxelDimensions", "Shininess", "HeightTextureSize"} <= names["VoxelConeTracing.fs"]
    u = uniforms.reference_uniforms()
    for f in names:
        for n in names[f] - {"DiffuseTexture", "SpecularTexture", "MaskTexture", "HeightTexture", "HeightTextureSize", "ShadowMap", "VoxelTexture",
                             "Shininess", "Opacity", "ShowDiffuse", "ShowIndirectDiffuse", "ShowSpecular", "ShowIndirectSpecular"}:
            assert n in u, (f, n)


def test_fly_camera_mirrors_camera_h():
    """Camera.h:21-25, 80-144: the constants and the movement rules of the fly camera, Python mirror and C++ facade."""
    from vct_b200 import renderer
    c = text("Camera.h")
    speed, sens = floats(r"const float SPEED = " + NUM, c, 1)[0], floats(r"const float SENSITIVITY = " + NUM, c, 1)[0]
    assert "MovementSpeed = SPEED;" in c and "MouseSensitivity = SENSITIVITY;" in c
    assert re.findall(r"^\t(\w+),?$", section(c, "enum Camera_Direction", "};"), flags=re.M) == ["FORWARD", "BACKWARD", "LEFT", "RIGHT", "UP", "DOWN"]
    cam = renderer.Camera()
    assert (cam.MovementSpeed, cam.MouseSensitivity) == (speed, sens) == (2.6, 0.1)
    assert [renderer.FORWARD, renderer.BACKWARD, renderer.LEFT, renderer.RIGHT, renderer.UP, renderer.DOWN] == list(range(6))
    np.testing.assert_allclose(cam.Front, [0, 0, -1], atol=1e-6)                       # Yaw = -90 looks down -z
    np.testing.assert_allclose(cam.Right, [1, 0, 0], atol=1e-6)
    p0 = cam.position.copy()
    cam.ProcessKeyBoard(renderer.FORWARD, 0.5); np.testing.assert_allclose(cam.position - p0, [0, 0, -1.3], atol=1e-6)
    cam.ProcessKeyBoard(renderer.RIGHT, 1.0); np.testing.assert_allclose(cam.position - p0, [2.6, 0, -1.3], atol=1e-6)
    cam.ProcessKeyBoard(renderer.DOWN, 1.0); np.testing.assert_allclose(cam.position - p0, [2.6, -2.6, -1.3], atol=1e-6)
    cam.ProcessMouseMovement(900.0, 2000.0)                                             # 90 degrees right, pitch clamps at 89
    assert cam.Yaw == 0.0 and cam.Pitch == 89.0 and "if (Pitch > 89.0f)" in c and "if (Pitch < -89.0f)" in c
    np.testing.assert_allclose(cam.Front, [np.cos(np.radians(89)), np.sin(np.radians(89)), 0], atol=1e-6)
    cam.ProcessMouseScroll(50.0); assert cam.Zoom == 1.0
    cam.ProcessMouseScroll(-100.0); assert cam.Zoom == 45.0 and "if (Zoom > 45.0f)" in c
    np.testing.assert_allclose(cam.GetViewMatrix() @ np.append(cam.position + cam.Front, 1.0), [0, 0, -1, 1], atol=1e-5)
    mine = open(os.path.join(ROOT, "voxel-cone-tracing_b200", "host", "Voxel_Cone_Tracing.h")).read()
    assert "MovementSpeed = 2.6f, MouseSensitivity = 0.1f" in mine
    for m in ("ProcessKeyBoard", "ProcessMouseMovement", "ProcessMouseScroll", "UpdateCamera", "GetViewMatrix"):
        assert re.search(rf"\b{m}\(", c) and re.search(rf"\b{m}\(", mine), m
