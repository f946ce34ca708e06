"""The Java sources under integration/ cannot be compiled here (no JDK in the image).  What can be checked without one:
they lex cleanly (Pygments' Java lexer finds no error token), brackets balance outside strings and comments, the package
matches the directory, the public type matches the file name, and every reference class they extend or import from
es.udc.fi.dc.irlab exists in the reference tree when that tree is present (it is not on the GPU box)."""
import os
import re

import pytest

pygments = pytest.importorskip("pygments")
from pygments.lexers import JavaLexer          # noqa: E402
from pygments.token import Token               # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA_ROOT = os.path.join(ROOT, "integration", "java")
REF_MAIN = "/root/reference/src/main/java"


def _sources():
    out = []
    for d, _, fs in os.walk(JAVA_ROOT):
        out += [os.path.join(d, f) for f in fs if f.endswith(".java")]
    return sorted(out)


def _top_level_args(text, start):
    """Number of arguments of the call whose '(' is at text[start]."""
    depth, n, seen = 0, 0, False
    for ch in text[start:]:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
            if depth == 0:
                return n + (1 if seen else 0)
        elif ch == "," and depth == 1:
            n += 1
        elif depth >= 1 and not ch.isspace():
            seen = True
    raise AssertionError("unterminated call")


def _check_static_uses(code, cls, ref_src):
    """Every `Cls.name(` / `Cls.NAME` our sources use exists in the reference class: static methods with an overload
    of the same number of parameters, constants by name."""
    ref_src = re.sub(r"/\*.*?\*/", "", ref_src, flags=re.S)
    for m in re.finditer(r"\b%s\.(\w+)\s*\(" % cls, code):
        name, n_args = m.group(1), _top_level_args(code, m.end() - 1)
        arities = set()
        for d in re.finditer(r"\bstatic\b[^;{=]*?\b%s\s*\(" % name, ref_src):
            arities.add(_top_level_args(ref_src, d.end() - 1))
        assert arities, "%s.%s() is not a static method of the reference class" % (cls, name)
        assert n_args in arities, "%s.%s called with %d arguments, the reference declares %s" % (cls, name, n_args, sorted(arities))
    for m in re.finditer(r"\b%s\.([A-Za-z_]\w*)\b(?!\s*\()" % cls, code):
        if m.group(1) == "class":
            continue
        assert re.search(r"\b%s\b" % m.group(1), ref_src), "%s.%s is not in the reference class" % (cls, m.group(1))


def test_there_are_java_sources():
    names = {os.path.basename(p) for p in _sources()}
    assert {"RM2Native.java", "RM2GpuJob.java", "AbstractRM2GpuReducer.java", "RM2GpuHDFSReducer.java",
            "RM2GpuCassandraReducer.java"} <= names


@pytest.mark.parametrize("path", _sources(), ids=lambda p: os.path.basename(p))
def test_java_source_is_well_formed(path):
    text = open(path, encoding="utf-8").read()
    toks = list(JavaLexer().get_tokens(text))
    assert not [v for t, v in toks if t in Token.Error], "lexer error tokens"
    pairs = {")": "(", "]": "[", "}": "{"}
    stack = []
    for t, v in toks:
        if t in Token.Comment or t in Token.Literal.String:
            continue
        for ch in v:
            if ch in "([{":
                stack.append(ch)
            elif ch in pairs:
                assert stack and stack.pop() == pairs[ch], "unbalanced %r" % ch
    assert not stack, "unclosed %r" % stack
    code = "".join(v for t, v in toks if t not in Token.Comment)
    # the reference's pom.xml compiles with <source>1.7</source>: no lambdas, no method references
    bare = "".join(v for t, v in toks if t not in Token.Comment and t not in Token.Literal.String)
    assert "->" not in bare and "::" not in bare, "Java 8 syntax in a -source 1.7 build"
    pkg = re.search(r"^\s*package\s+([\w.]+)\s*;", code, re.M)
    assert pkg, "no package declaration"
    rel = os.path.relpath(os.path.dirname(path), JAVA_ROOT).replace(os.sep, ".")
    assert pkg.group(1) == rel
    base = os.path.splitext(os.path.basename(path))[0]
    if base != "package-info":
        assert re.search(r"\b(class|interface|enum)\s+%s\b" % re.escape(base), code), "type name differs from the file name"
    if os.path.isdir(REF_MAIN):
        ours = {os.path.splitext(os.path.relpath(p, JAVA_ROOT))[0].replace(os.sep, ".") for p in _sources()}
        for imp in re.findall(r"^\s*import\s+(es\.udc\.fi\.dc\.irlab\.[\w.]+)\s*;", code, re.M):
            if imp in ours:
                continue
            ref = os.path.join(REF_MAIN, imp.replace(".", os.sep) + ".java")
            assert os.path.exists(ref), "import %s not in the reference" % imp
            _check_static_uses(code, imp.split(".")[-1], open(ref, encoding="utf-8").read())
