"""ORACLE — TEST INFRASTRUCTURE ONLY.

gen.py — build step of oracle/_ref/libhelios_glsl_ref.so: reads the reference's shader files WHERE THEY LIE
(/root/reference/src/engine/shader) and writes g++-compilable copies into oracle/_ref/gen/ (git-ignored build
output; reference sources are never committed).  The rewrite is purely lexical — no statement of any shader
function is touched; only the things C++ has no syntax for are mapped:

  1. `#version` / `#extension` lines                          -> dropped
  2. `#include "x"`                                           -> `#include "x.inc"`
  3. unsuffixed floating literals (`2.0`, `0.99`, `1.0/2.2`)  -> `2.0f` ... (GLSL literals are 32-bit floats)
  4. parameter qualifiers: `in T a` -> `T a`; `out T a` / `inout T a` -> `T& a`
  5. interface blocks `layout(..) [readonly|writeonly] buffer Name { T data[]; } Inst[];`
                                                              -> `struct Name { T* data; }; static Name* Inst;`
     `layout(..) uniform Name { .. } inst;` (UBO / push consts) -> `struct Name { .. }; static Name inst;`
  6. opaque uniforms `layout(..) uniform sampler2D s[];` etc.  -> `static sampler2D_array s;` / `static T name;`
  7. `rayPayloadEXT` / `rayPayloadInEXT` / `hitAttributeEXT` / stage `in` / `out` variables
                                                              -> `static thread_local T name;`
  8. constructor calls `vec2(a, b)`, `vec3(..)`, `uvec2(..)`, `mat3(..)` ... -> `vec2{a, b}`: GLSL evaluates
     arguments left to right (sampling.glsl:20 draws two random numbers inside one constructor); C++ guarantees
     that order only for braced lists
The host half of the sky model (C++ in the reference) is cut out of gfx/hosek_wilkie_sky_model.cpp by function
name, unmodified (extract_sky_host below), because the rest of that file is Vulkan object handling.
Everything else (swizzles, constructors, built-ins, traceRayEXT, ignoreIntersectionEXT, main) is handled by
glsl_compat.h and by macros in the stage_*.cpp wrappers.

    python oracle/ref_glsl/gen.py /root/reference/src/engine/shader oracle/_ref/gen
"""
from __future__ import annotations

import re
import sys
from pathlib import Path

FILES = [
    "random.glsl", "sampling.glsl", "common.glsl", "brdf.glsl",
    "path_trace_rgen.glsl", "path_trace_rchit.glsl", "path_trace_rahit.glsl", "path_trace_rmiss.glsl",
    "path_trace_shadow.rchit", "path_trace_shadow.rmiss", "tone_map.frag", "procedural_sky.frag",
]

_FLOAT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+)(?![\w.])")
_LAYOUT = r"layout\s*\([^)]*\)\s*"


def _block(m: re.Match) -> str:
    name, body, inst, arr = m.group("name"), m.group("body"), m.group("inst"), m.group("arr")
    body = re.sub(r"(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"\1* \2;", body)  # runtime-sized array member -> pointer
    decl = f"static {name}* {inst};" if arr else f"static {name} {inst};"
    return f"struct {name}\n{{{body}}};\n{decl}"


_CTOR = re.compile(r"\b(vec[234]|uvec[234]|ivec[234]|mat3|mat4)\(")


def _brace_constructors(text: str) -> str:
    out, i = [], 0
    while True:
        m = _CTOR.search(text, i)
        if not m:
            out.append(text[i:])
            return "".join(out)
        # a declaration such as `vec3 f(...)` never has the type directly before `(`, so this is a constructor call
        depth, j = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(text[j], 0)
            j += 1
        inner = _brace_constructors(text[m.end():j - 1])
        out.append(text[i:m.start()] + m.group(1) + "{" + inner + "}")
        i = j


def rewrite(text: str) -> str:
    text = re.sub(r"^[ \t]*#(version|extension)[^\n]*\n", "", text, flags=re.M)
    text = re.sub(r'#include\s+"([^"]+)"', r'#include "\1.inc"', text)
    text = _FLOAT.sub(r"\1f", text)
    # interface blocks (buffers, UBOs, push constants)
    text = re.sub(
        _LAYOUT + r"(?:readonly\s+|writeonly\s+)?(?:buffer|uniform)\s+(?P<name>\w+)\s*\{(?P<body>[^}]*)\}\s*(?P<inst>\w+)\s*(?P<arr>\[\s*\])?\s*;",
        _block, text)
    # opaque uniforms
    text = re.sub(_LAYOUT + r"(?:readonly\s+|writeonly\s+)?uniform\s+(\w+)\s+(\w+)\s*\[\s*\]\s*;", r"static \1_array \2;", text)
    text = re.sub(_LAYOUT + r"(?:readonly\s+|writeonly\s+)?uniform\s+(\w+)\s+(\w+)\s*;", r"static \1 \2;", text)
    # payloads, hit attributes, stage inputs / outputs
    text = re.sub(r"(?:" + _LAYOUT + r")?\b(?:rayPayloadInEXT|rayPayloadEXT|hitAttributeEXT)\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text)
    text = re.sub(_LAYOUT + r"(?:in|out)\s+(\w+)\s+(\w+)\s*;", r"static thread_local \1 \2;", text)
    # parameter qualifiers
    text = re.sub(r"([(,]\s*)(?:inout|out)\s+(\w+)\s+(\w+)", r"\1\2& \3", text)
    text = re.sub(r"([(,]\s*)in\s+(?=\w+\s+\w+)", r"\1", text)
    text = _brace_constructors(text)
    if re.search(r"\blayout\s*\(", text):
        raise SystemExit("gen.py: a layout(...) declaration was not recognised:\n" + "\n".join(l for l in text.splitlines() if "layout" in l))
    return text


def _balanced(text: str, start: int) -> int:
    """index just past the `}` that closes the first `{` at or after `start`"""
    j = text.index("{", start)
    depth = 0
    while True:
        depth += {"{": 1, "}": -1}.get(text[j], 0)
        j += 1
        if depth == 0:
            return j


def extract_sky_host(cpp: str) -> tuple[str, str]:
    """The host half of the Hosek-Wilkie model (src/engine/gfx/hosek_wilkie_sky_model.cpp of the reference): the three
    free functions evaluate_spline / evaluate / hosek_wilkie verbatim, and the coefficient statements of
    HosekWilkieSkyModel::update (from `const float sunTheta` to the end of the `if (m_normalized_sun_y)` block) —
    everything between and after them is Vulkan object handling and is not on the path."""
    funcs = []
    for head in ("double evaluate_spline(", "double evaluate(", "glm::vec3 hosek_wilkie("):
        a = cpp.index(head)
        funcs.append(cpp[a:_balanced(cpp, a)])
    u = cpp.index("HosekWilkieSkyModel::update(")
    a = cpp.index("const float sunTheta", u)
    b = _balanced(cpp, cpp.index("if (m_normalized_sun_y)", a))
    return "\n\n".join(funcs) + "\n", cpp[a:b] + "\n"


def main(src: str, dst: str) -> None:
    s, d = Path(src), Path(dst)
    d.mkdir(parents=True, exist_ok=True)
    for f in FILES:
        out = rewrite((s / f).read_text())
        (d / (f + ".inc")).write_text("// GENERATED by oracle/ref_glsl/gen.py from the reference's " + f + " — build output, do not commit\n" + out)
    sky_cpp = s.parent / "gfx" / "hosek_wilkie_sky_model.cpp"
    funcs, body = extract_sky_host(sky_cpp.read_text())
    note = "// GENERATED by oracle/ref_glsl/gen.py from the reference's gfx/hosek_wilkie_sky_model.cpp — build output, do not commit\n"
    (d / "hosek_host_functions.inc").write_text(note + funcs)
    (d / "hosek_host_update.inc").write_text(note + body)
    print(f"gen.py: {len(FILES)} shader files + the Hosek-Wilkie host functions -> {d}")


if __name__ == "__main__":
    main(*sys.argv[1:3])
