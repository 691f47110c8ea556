#!/usr/bin/env python
"""Generate rust/searchlite-gpu-sys/src/lib.rs — the `extern "C"` block and #[repr(C)] types a searchlite-core
`gpu` feature would bind — from include/searchlite_gpu.h.

The header is written in a regular subset of C (typedef'd anonymous structs of scalar / pointer fields, anonymous and
typedef'd enums, `#define NAME <int>u`, one prototype per `;`), which this script parses directly: no bindgen, no
libclang (neither exists in this image, and neither does a Rust toolchain — the crate is checked for being in sync with
the header by tests/test_abi_and_host.py, not compiled here).

Usage: python tools/gen_rust_sys.py [--check]
"""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "searchlite_gpu.h")
OUT = os.path.join(ROOT, "rust", "searchlite-gpu-sys", "src", "lib.rs")

SCALARS = {
    "uint8_t": "u8", "int8_t": "i8", "uint16_t": "u16", "int16_t": "i16", "uint32_t": "u32", "int32_t": "i32",
    "uint64_t": "u64", "int64_t": "i64", "float": "f32", "double": "f64", "char": "c_char", "void": "c_void", "size_t": "usize",
}


def strip_comments(src: str) -> str:
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def rust_type(ctype: str, enums: set, structs: set) -> str:
    """`const uint32_t *const *` -> `*const *const u32`"""
    t = ctype.strip()
    ptrs = []
    while True:
        t = t.strip()
        m = re.match(r"^(.*)\*\s*(const)?$", t)
        if not m:
            break
        t = m.group(1)
        ptrs.append("const" if m.group(2) else None)  # constness of the POINTER itself: irrelevant in Rust's raw pointer type
    base_const = bool(re.search(r"\bconst\b", t))
    base = re.sub(r"\bconst\b|\bstruct\b", " ", t).strip()
    if base in SCALARS:
        r = SCALARS[base]
    elif base in enums:
        r = "i32" if base in ("slg_status",) else "u32"
        r = base  # typedef'd enums become type aliases below
    elif base in structs or base in ("slg_index_t", "slg_batch_t"):
        r = base
    else:
        raise ValueError(f"unknown C type {ctype!r}")
    # innermost pointer first: its pointee constness is the base's; outer pointers point at (possibly const) pointers
    n = len(ptrs)
    for i in range(n):
        pointee_const = base_const if i == 0 else (ptrs[n - i] == "const")
        r = ("*const " if pointee_const else "*mut ") + r
    return r


def parse(src: str):
    src = strip_comments(src)
    defines = re.findall(r"#define\s+(SLG_\w+)\s+(\d+)u?\b", src)
    enums_named, consts = {}, []
    for m in re.finditer(r"(typedef\s+)?enum\s*\{(.*?)\}\s*(\w+)?\s*;", src, flags=re.S):
        name = m.group(3)
        items = []
        nxt = 0
        for it in m.group(2).split(","):
            it = it.strip()
            if not it:
                continue
            if "=" in it:
                k, v = [x.strip() for x in it.split("=")]
                val = int(v.rstrip("u"), 0)
            else:
                k, val = it, nxt
            nxt = val + 1
            items.append((k, val))
        if name:
            enums_named[name] = items
        else:
            consts += items
    structs = {}
    for m in re.finditer(r"typedef\s+struct\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        fields = []
        for decl in m.group(1).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            # `int64_t i_min, i_max` / `const uint32_t *post_docs` / `char name[64]`
            first, *rest = [d.strip() for d in decl.split(",")]
            fm = re.match(r"^(.*?)(\w+)(\[\d+\])?$", first)
            ctype = fm.group(1)
            names = [(fm.group(2), fm.group(3))]
            base_type = re.sub(r"\*", " ", ctype).strip()
            for r_ in rest:
                rm = re.match(r"^(\*?)\s*(\w+)(\[\d+\])?$", r_)
                names.append((rm.group(2), rm.group(3)))
                assert not rm.group(1), "mixed pointer declarators are not used in the header"
            for nm, arr in names:
                fields.append((nm, ctype.strip(), arr, base_type))
        structs[m.group(2)] = fields
    protos = []
    body = src[src.index("typedef struct slg_index slg_index_t;"):]
    body = re.sub(r"typedef\s+struct\s*\{.*?\}\s*\w+\s*;", " ", body, flags=re.S)
    body = re.sub(r"(typedef\s+)?enum\s*\{.*?\}\s*\w*\s*;", " ", body, flags=re.S)
    body = re.sub(r"^\s*#[^\n]*$", ";", body, flags=re.M)  # preprocessor lines end a statement
    body = re.sub(r'extern\s+"C"\s*\{', ";", body)
    for m in re.finditer(r"([\w\s\*]+?)\b(slg_\w+)\s*\(([^;{}]*?)\)\s*;", body, flags=re.S):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        if "typedef" in ret:
            continue
        params = []
        if args and args != "void":
            for i, a in enumerate(args.split(",")):
                a = a.strip()
                am = re.match(r"^(.*?)(\w+)?$", a)
                # a parameter is `type name` or just `type` (handles: `slg_index_t *`)
                if am.group(2) and am.group(2) not in SCALARS and not am.group(2).endswith("_t") and am.group(2) != "const" and am.group(1).strip():
                    params.append((am.group(2), am.group(1).strip()))
                else:
                    params.append((f"arg{i}", a))
        protos.append((name, ret, params))
    return defines, enums_named, consts, structs, protos


RUST_KEYWORDS = {"type", "match", "ref", "in", "fn", "mod", "use", "box", "move", "loop", "where", "impl", "dyn", "as", "self", "super"}


def ident(n: str) -> str:
    return n + "_" if n in RUST_KEYWORDS else n


def generate() -> str:
    defines, enums_named, consts, structs, protos = parse(open(HEADER).read())
    enum_names, struct_names = set(enums_named), set(structs)
    out = []
    w = out.append
    w("// GENERATED by tools/gen_rust_sys.py from include/searchlite_gpu.h — do not edit by hand.")
    w("//")
    w("// Raw FFI surface of libsearchlite_gpu.so for searchlite-core's `gpu` feature (SURVEY.md §8b).  Semantics, ownership")
    w("// and the reference lines each entry point replaces are documented in the header; INTEGRATION.md shows the call site.")
    w("#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]")
    w("")
    w("use core::ffi::{c_char, c_void};")
    w("")
    w("/// opaque: owns the device-resident segments of one index (SegmentReader::open's replacement)")
    w("#[repr(C)]")
    w("pub struct slg_index_t {")
    w("    _private: [u8; 0],")
    w("}")
    w("/// opaque: one prepared query batch")
    w("#[repr(C)]")
    w("pub struct slg_batch_t {")
    w("    _private: [u8; 0],")
    w("}")
    w("")
    for name, val in defines:
        w(f"pub const {name}: u32 = {val};")
    w("")
    for name, items in enums_named.items():
        signed = any(v < 0 for _, v in items)
        w(f"pub type {name} = {'i32' if signed else 'u32'};")
        for k, v in items:
            w(f"pub const {k}: {name} = {v};")
        w("")
    for k, v in consts:
        w(f"pub const {k}: u32 = {v};")
    w("")
    for name, fields in structs.items():
        w("#[repr(C)]")
        w("#[derive(Clone, Copy, Debug)]")
        w(f"pub struct {name} {{")
        for nm, ctype, arr, _ in fields:
            rt = rust_type(ctype, enum_names, struct_names)
            if arr:
                rt = f"[{rt}; {int(arr[1:-1])}]"
            w(f"    pub {ident(nm)}: {rt},")
        w("}")
        w("")
    w('#[link(name = "searchlite_gpu")]')
    w('extern "C" {')
    for name, ret, params in protos:
        ps = ", ".join(f"{ident(pn)}: {rust_type(pt, enum_names, struct_names)}" for pn, pt in params)
        rr = rust_type(ret, enum_names, struct_names)
        w(f"    pub fn {name}({ps}) -> {rr};")
    w("}")
    w("")
    return "\n".join(out)


def exported_names():
    return [p[0] for p in parse(open(HEADER).read())[4]]


def main() -> int:
    text = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            sys.stderr.write("rust/searchlite-gpu-sys/src/lib.rs is out of date: run python tools/gen_rust_sys.py\n")
            return 1
        return 0
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write(text)
    print(OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
