#!/usr/bin/env python3
"""Emit the consts.cu the reference's build would generate (vision-common/build.rs:1-15 +
toml-consts/src/lib.rs:31-59,185-197) from the reference's own consts.toml, so that
vision-gpu/cuda/cuda.cu compiles unmodified as the GPU comparator.  Output goes to oracle/_ref/.

usage: gen_ref_consts.py /root/reference out/consts.cu
"""
import sys
import tomllib
from pathlib import Path

CTYPE = {"u8": "uint8_t", "i8": "int8_t", "u16": "uint16_t", "i16": "int16_t", "u32": "uint32_t",
         "i32": "int32_t", "u64": "uint64_t", "i64": "int64_t", "f32": "float", "f64": "double"}


def fmt(v, ty):
    if ty in ("f32", "f64"):
        return repr(float(v))
    return str(int(v))


def main(ref: str, out: str) -> None:
    toml = tomllib.loads((Path(ref) / "vision-common/src/consts/consts.toml").read_text())
    lines = ["#define SMH_CONSTS", ""]
    for name in sorted(toml):  # BTreeMap order
        ty, val = toml[name]["type"], toml[name]["value"]
        c = CTYPE[ty]
        if isinstance(val, list):
            lines.append(f"constexpr __device__ {c} {name}[{len(val)}] = {{{', '.join(fmt(v, ty) for v in val)}}};")
        else:
            lines.append(f"constexpr __device__ {c} {name} = {fmt(val, ty)};")
    Path(out).parent.mkdir(parents=True, exist_ok=True)
    Path(out).write_text("\n".join(lines) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
