"""The reference-side binding files under scala/ cannot be compiled here (no JDK / scalac), so they are kept honest
mechanically: the JavaCPP preset declares exactly the header's symbols, every gs_* call in the Scala shims exists in the
preset with the right number of arguments, and tests/golden/jvm_expected.txt is what the oracle produces today."""
import importlib.util
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALA = os.path.join(ROOT, "scala")


def _gen():
    spec = importlib.util.spec_from_file_location("gen_javacpp", os.path.join(ROOT, "tools", "gen_javacpp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_preset_is_the_generated_one_and_covers_every_symbol():
    gen = _gen()
    with open(gen.OUT) as f:
        committed = f.read()
    assert committed == gen.emit(), "run python tools/gen_javacpp.py after changing include/gs_b200.h"
    declared = {name for _, name, _ in gen.prototypes()}
    native = set(re.findall(r"native [^;]*?\b(gs_\w+)\(", committed))
    assert declared == native and len(declared) >= 60


def test_scala_shims_call_existing_symbols_with_the_right_arity():
    gen = _gen()
    arity = {name: len(args) for _, name, args in gen.prototypes()}
    for fn in ("Flatten.scala", "B200OneDGrid.scala", "B200TwoDGrid.scala"):
        src = open(os.path.join(SCALA, "approximation", "b200", fn)).read()
        src = re.sub(r"//.*", "", src)
        for m in re.finditer(r"\b(gs_\w+)\(", src):
            name = m.group(1)
            if name in ("gs_ctx", "gs_grid1d", "gs_grid2d", "gs_ragged"):
                continue                      # handle constructors
            assert name in arity, f"{fn}: {name} is not in include/gs_b200.h"
            # count top-level commas of the call
            depth, i, commas, empty = 0, m.end(), 0, True
            while True:
                ch = src[i]
                if ch in "([":
                    depth += 1
                elif ch in ")]":
                    if depth == 0:
                        break
                    depth -= 1
                elif ch == "," and depth == 0:
                    commas += 1
                if not ch.isspace() and not (ch == ")" and depth == 0):
                    empty = False
                i += 1
            got = 0 if empty else commas + 1
            assert got == arity[name], f"{fn}: {name} called with {got} arguments, the header declares {arity[name]}"


def test_jvm_expected_is_what_the_oracle_produces():
    path = os.path.join(ROOT, "tests", "golden", "jvm_expected.txt")
    before = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "golden", "make_jvm_expected.py")],
                          stdout=subprocess.DEVNULL)
    assert open(path).read() == before
    names = [ln.split()[0] for ln in before.splitlines()]
    src = open(os.path.join(SCALA, "OracleCheck.scala")).read()
    for n in names:
        stem = re.sub(r"_\d+$", "", n)
        assert stem in src, f"OracleCheck.scala does not print {n}"
