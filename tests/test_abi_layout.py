"""CPU: the ctypes mirrors in colpali_b200/_lib.py have the layout a C compiler gives the structs of
include/colpali_b200.h -- field by field (name, order, offset, size) and in total.  A probe program is generated from the
header's struct bodies, compiled with the C compiler (the header must stay valid plain C, not just C++) and its
offsetof / sizeof output compared with ctypes.  A mismatch here is silent memory corruption on the GPU path."""
import ctypes
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT
from colpali_b200 import _lib

HEADER = os.path.join(ROOT, "include", "colpali_b200.h")
MIRRORS = {
    "cpb_loss_desc": _lib.LossDesc,
    "cpb_maxsim_args": _lib.MaxSimArgs,
    "cpb_maxsim_bwd_args": _lib.MaxSimBwdArgs,
    "cpb_exchange_push_args": _lib.ExchangePushArgs,
    "cpb_dense_dot_args": _lib.DenseDotArgs,
}


def header_structs():
    """{struct name: [field names in declaration order]} parsed from the header (comments stripped)."""
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s*\{(.*?)\}\s*(\w+)\s*;", src, flags=re.S):
        assert m.group(1) == m.group(3)
        fields = []
        for decl in m.group(2).split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for piece in decl.split(","):  # "int32_t a, b" declares two members
                name = re.search(r"(\w+)\s*(\[\s*\w+\s*\])?$", piece.strip())
                assert name, decl
                fields.append(name.group(1))
        out[m.group(1)] = fields
    return out


def test_every_header_struct_has_a_mirror_and_vice_versa():
    assert sorted(header_structs()) == sorted(MIRRORS)


def test_field_names_and_order_match_the_header():
    for name, fields in header_structs().items():
        assert [f[0] for f in MIRRORS[name]._fields_] == fields, name


def test_offsets_and_sizes_match_what_the_c_compiler_lays_out(tmp_path):
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = header_structs()
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for name, fields in structs.items():
        lines.append(f'  printf("{name} . %zu %zu\\n", sizeof({name}), (size_t)_Alignof({name}));')
        for f in fields:
            lines.append(f'  printf("{name} {f} %zu %zu\\n", offsetof({name}, {f}), sizeof((({name}*)0)->{f}));')
    lines += ['  printf("abi . %d 0\\n", CPB_ABI_VERSION);', "  return 0;", "}"]
    src, exe = tmp_path / "probe.c", tmp_path / "probe"
    src.write_text("\n".join(lines))
    subprocess.run([cc, "-std=c11", "-Wall", "-Werror", "-o", str(exe), str(src)], check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        name, field, a, b = line.split()
        if name == "abi":
            assert int(a) == _lib.CPB_ABI_VERSION
            continue
        mirror = MIRRORS[name]
        if field == ".":
            assert ctypes.sizeof(mirror) == int(a), f"sizeof({name}): ctypes {ctypes.sizeof(mirror)} vs C {a}"
            assert ctypes.alignment(mirror) == int(b), name
        else:
            desc = getattr(mirror, field)
            assert (desc.offset, desc.size) == (int(a), int(b)), f"{name}.{field}: ctypes {(desc.offset, desc.size)} vs C {(a, b)}"
            seen += 1
    assert seen == sum(len(f) for f in structs.values())


def test_constants_match_the_header():
    src = open(HEADER).read()
    for c_name, value in re.findall(r"#define\s+(CPB_[A-Z0-9_]+)\s+\(?(-?\d+)u?\)?\s", src):
        if hasattr(_lib, c_name):
            assert getattr(_lib, c_name) == int(value), c_name
    # the ones the host code relies on must exist on both sides
    for c_name in ("CPB_ABI_VERSION", "CPB_FLAG_ROUND_BF16", "CPB_FLAG_CONTIGUOUS", "CPB_FLAG_INDEPENDENT", "CPB_FLAG_GRAD_BF16",
                   "CPB_LOSS_CE", "CPB_LOSS_PAIRWISE", "CPB_LOSS_SIGMOID", "CPB_LOSS_SYMMETRIC_CE", "CPB_TOPK_MAX"):
        assert re.search(rf"#define\s+{c_name}\b", src), c_name
        assert hasattr(_lib, c_name), c_name
