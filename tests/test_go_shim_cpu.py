"""Static consistency of the cgo shim (go/, source only: no Go toolchain in this image) with the
C ABI it binds: every C.gopf_* call names a function include/gopf_cuda.h declares and passes the
declared number of arguments; every C type it mentions is declared.  No GPU needed."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def strip_c_comments(src: str) -> str:
    return re.sub(r"/\*.*?\*/", "", src, flags=re.S)


def split_top_level(args: str):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def header_prototypes():
    src = strip_c_comments(open(os.path.join(ROOT, "include", "gopf_cuda.h")).read())
    protos = {}
    for m in re.finditer(r"\b(gopf_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        if name == "gopf_time_fn":
            continue
        n = 0 if args in ("", "void") else len(split_top_level(args))
        protos[name] = n
    return protos


def go_calls():
    calls = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "go")):
        for fn in files:
            if not fn.endswith(".go"):
                continue
            src = open(os.path.join(dirpath, fn)).read()
            src = re.sub(r"//[^\n]*", "", src)
            # drop the cgo preamble (C code, not Go calls)
            src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
            for m in re.finditer(r"C\.(gopf_[a-z0-9_]+)\s*\(", src):
                i = m.end()
                depth, j = 1, i
                while depth and j < len(src):
                    depth += src[j] in "([{"
                    depth -= src[j] in ")]}"
                    j += 1
                args = src[i:j - 1].strip()
                calls.append((fn, m.group(1), 0 if not args else len(split_top_level(args))))
    return calls


def test_every_cgo_call_matches_a_declared_prototype():
    protos = header_prototypes()
    assert len(protos) >= 60
    calls = go_calls()
    assert len(calls) >= 30
    helpers = {"gopf_source_trampoline_ptr": 0, "gopf_index_as_ptr": 1}  # static helpers of the cgo preamble
    bad = []
    for fn, name, nargs in calls:
        want = protos.get(name, helpers.get(name))
        if want is None:
            bad.append(f"{fn}: C.{name} is not declared in include/gopf_cuda.h")
        elif want != nargs:
            bad.append(f"{fn}: C.{name} called with {nargs} arguments, declared with {want}")
    assert not bad, "\n".join(bad)


def test_cgo_types_are_declared():
    hdr = strip_c_comments(open(os.path.join(ROOT, "include", "gopf_cuda.h")).read())
    for dirpath, _, files in os.walk(os.path.join(ROOT, "go")):
        for fn in files:
            if fn.endswith(".go"):
                src = open(os.path.join(dirpath, fn)).read()
                for t in set(re.findall(r"C\.(gopf_[a-z0-9_]+)\b(?!\s*\()", src)):
                    assert re.search(r"\b" + t + r"\b", hdr), f"{fn}: C.{t} unknown"
