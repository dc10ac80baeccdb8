"""Extract, VERBATIM, the reference's own source of the hot path into oracle/_ref/gen/*.inc (git-ignored).

TEST INFRASTRUCTURE.  The reference's headers on this path (`Shape.hpp`, `sw_manager.hpp`, `back_end_optimizer.hpp`)
cannot be compiled as whole files here: they pull in ROS, PCL, libigl-with-Eigen, yaml and visualisation members that
have nothing to do with the arithmetic.  What CAN be compiled is the arithmetic itself — the shape classes, the
SweptVolumeManager query methods and the TrajOptimizer penalty/chain-rule methods are plain C++ over a small Eigen
surface.  This script cuts exactly those definitions out of the files where they lie under /root/reference, byte for
byte (class bodies / member functions located by their signature and brace matching — no edits), and writes them to
include fragments; `oracle/ref_path_shim.cpp` (ours) supplies only the scaffolding around them (a stub `BasicShape`
base with `trans/Rotate/getTransform`, a `SweptVolumeManager` and `TrajOptimizer` holding the members the methods
touch) and `oracle/ref_shim/Eigen/*` supplies the Eigen surface.  `trajectory.hpp` and `minco.hpp` are NOT extracted:
they are included whole from the reference tree.

Nothing generated here is committed (oracle/_ref/ is git-ignored): reference sources never enter the repository.

    python oracle/ref_extract.py [--ref /root/reference] [--out oracle/_ref/gen]
"""
from __future__ import annotations

import argparse
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _skip_noncode(s: str, i: int) -> int:
    """If s[i:] starts a comment / string / char literal return the index just past it, else i."""
    if s.startswith("//", i):
        j = s.find("\n", i)
        return len(s) if j < 0 else j
    if s.startswith("/*", i):
        j = s.find("*/", i + 2)
        return len(s) if j < 0 else j + 2
    if s[i] == '"':
        j = i + 1
        while j < len(s) and s[j] != '"':
            j += 2 if s[j] == "\\" else 1
        return j + 1
    if s[i] == "'":
        j = i + 1
        while j < len(s) and s[j] != "'":
            j += 2 if s[j] == "\\" else 1
        return j + 1
    return i


def match_braces(s: str, start: int) -> int:
    """start: index of an opening '{'.  Returns the index just past its matching '}'."""
    assert s[start] == "{"
    depth = 0
    i = start
    while i < len(s):
        j = _skip_noncode(s, i)
        if j != i:
            i = j
            continue
        c = s[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced braces")


def line_of(s: str, idx: int) -> int:
    return s.count("\n", 0, idx) + 1


class Source:
    def __init__(self, path: str):
        self.path = path
        with open(path, encoding="utf-8", errors="surrogateescape") as fh:
            self.text = fh.read()

    def block(self, pattern: str, occurrence: int = 0, expect_lines=None, trailing_semicolon=False) -> str:
        """The definition whose first line matches `pattern` (regex, MULTILINE), `occurrence`-th match, from the start of
        that line (including a directly preceding `template <...>` line) to the matching closing brace."""
        ms = list(re.finditer(pattern, self.text, re.M))
        if len(ms) <= occurrence:
            raise KeyError(f"{self.path}: pattern {pattern!r} occurrence {occurrence} not found ({len(ms)} matches)")
        m = ms[occurrence]
        start = self.text.rfind("\n", 0, m.start()) + 1
        # pull in a template header on the previous non-blank line
        prev_end = start - 1
        prev_start = self.text.rfind("\n", 0, prev_end) + 1
        if re.match(r"\s*template\s*<", self.text[prev_start:prev_end]):
            start = prev_start
        brace = self.text.index("{", m.end() - 1 if self.text[m.end() - 1] == "{" else m.end())
        end = match_braces(self.text, brace)
        if trailing_semicolon:
            k = end
            while self.text[k] in " \t\r\n":
                k += 1
            # `} Name;` (typedef class ... Name;) or `};`
            semi = self.text.index(";", k)
            assert semi - k < 40, (pattern, self.text[k:semi])
            end = semi + 1
        l0, l1 = line_of(self.text, start), line_of(self.text, end - 1)
        if expect_lines is not None:
            lo, hi = expect_lines
            if not (lo <= l0 and l1 <= hi):
                raise AssertionError(f"{self.path}: {pattern!r} found at lines {l0}-{l1}, expected within {lo}-{hi} "
                                     "(different reference revision?)")
        rel = os.path.relpath(self.path, "/root/reference")
        return f"// ---- verbatim {rel}:{l0}-{l1}\n#line {l0} \"{self.path}\"\n" + self.text[start:end] + "\n"

    def lines(self, first_pattern: str, last_pattern: str, expect_lines=None) -> str:
        """Whole lines from the first line matching `first_pattern` to the next line matching `last_pattern` (inclusive)."""
        m0 = re.search(first_pattern, self.text, re.M)
        if not m0:
            raise KeyError(first_pattern)
        m1 = re.compile(last_pattern, re.M).search(self.text, m0.start())
        if not m1:
            raise KeyError(last_pattern)
        start = self.text.rfind("\n", 0, m0.start()) + 1
        end = self.text.find("\n", m1.end())
        l0, l1 = line_of(self.text, start), line_of(self.text, end - 1)
        if expect_lines is not None and not (expect_lines[0] <= l0 and l1 <= expect_lines[1]):
            raise AssertionError(f"{self.path}: {first_pattern!r} found at lines {l0}-{l1}, expected within {expect_lines}")
        rel = os.path.relpath(self.path, "/root/reference")
        return f"// ---- verbatim {rel}:{l0}-{l1}\n#line {l0} \"{self.path}\"\n" + self.text[start:end] + "\n"

    def macro(self, name: str) -> str:
        m = re.search(r"^#define\s+" + re.escape(name) + r"\b.*$", self.text, re.M)
        if not m:
            raise KeyError(name)
        end = m.end()
        while self.text[end - 1] == "\\" or self.text[m.start():end].rstrip().endswith("\\"):
            nxt = self.text.find("\n", end + 1)
            end = len(self.text) if nxt < 0 else nxt
            if not self.text[:end].rstrip(" \t").endswith("\\"):
                break
        l0, l1 = line_of(self.text, m.start()), line_of(self.text, end)
        rel = os.path.relpath(self.path, "/root/reference")
        return f"// ---- verbatim {rel}:{l0}-{l1}\n" + self.text[m.start():end] + "\n"


SHAPE_CLASSES = ["Circle", "sdUnevenCapsule", "star", "sdTunnel", "sdCutDisk", "sdTrapezoid", "sdRhombus", "sdHorseshoe",
                 "sdHeart", "sdRoundedX", "bigX", "sdRoundedCross", "sdOrientedVesica", "sdMoon", "sdPie", "sdPie2",
                 "sdArc", "Polygon"]


def extract(ref: str, out: str) -> None:
    os.makedirs(out, exist_ok=True)
    inc = os.path.join(ref, "src")
    shape = Source(os.path.join(inc, "utils/include/utils/Shape.hpp"))
    swm = Source(os.path.join(inc, "swept_volume/include/swept_volume/sw_manager.hpp"))
    beo = Source(os.path.join(inc, "planner_algorithm/include/planner_algorithm/back_end_optimizer.hpp"))

    files = {}
    # --- Shape.hpp -------------------------------------------------------------------------------------------------
    files["shape_macros.inc"] = shape.macro("PI") + shape.macro("DEFINE_USEFUL_FUNCTION")
    files["shape_or_mask.inc"] = shape.lines(r"^\s*const uint8_t or_mask\[8\]", r";", expect_lines=(90, 100))
    # the part of BasicShape::BasicShape that sets the body-frame pre-transform from poly_params (the rest loads the mesh)
    files["shape_base_ctor_transform.inc"] = shape.lines(r"^\s*trans = Eigen::Vector3d\(para\[0\], para\[1\], 0\.0\);",
                                                         r"^\s*Rotate\(2, 2\) = 1;", expect_lines=(284, 296))
    files["shape_base_kernels.inc"] = (
        shape.block(r"^\s*typedef class shapeKernel\s*$", expect_lines=(95, 150), trailing_semicolon=True)
        + shape.block(r"^\s*typedef class byteShapeKernel\s*$", expect_lines=(145, 225), trailing_semicolon=True))
    files["shape_base_initshape.inc"] = shape.block(r"^\s*void initShape\(const double ndx", expect_lines=(380, 435))
    files["shape_classes.inc"] = "".join(
        shape.block(r"^\s*class " + c + r" : public BasicShape\s*$", trailing_semicolon=True) for c in SHAPE_CLASSES)

    # --- sw_manager.hpp --------------------------------------------------------------------------------------------
    files["sw_macros.inc"] = swm.macro("TRAJ_ORDER") + swm.macro("useScale") + swm.macro("useNumer") + swm.macro("pi")
    files["sw_sampleset.inc"] = (swm.block(r"^class CircleCoord2D\s*$", expect_lines=(20, 45), trailing_semicolon=True)
                                 + swm.block(r"^class SampleSet2D\s*$", expect_lines=(38, 130), trailing_semicolon=True))
    sig = r"^\s*inline void getStateOnTrajStamp\(const double &time_stamp,\s*$"
    methods = [
        swm.block(r"^\s*inline void updateTraj\(", expect_lines=(370, 390)),
        swm.block(sig, 0, expect_lines=(410, 440)),  # (t, xt, vt, Rt, VRt)
        swm.block(sig, 1, expect_lines=(435, 465)),  # (t, xt, vt, Rt, VRt, St, dSt)
        swm.block(sig, 2, expect_lines=(460, 480)),  # (t, xt, Rt)   <- the one on the path
        swm.block(sig, 3, expect_lines=(470, 495)),  # (t, xt, Rt, St)
        swm.block(r"^\s*inline Eigen::Matrix3d getScale\(", expect_lines=(485, 510)),
        swm.block(r"^\s*inline Eigen::Matrix3d getDotScale\(", expect_lines=(500, 525)),
        swm.block(r"^\s*inline Eigen::Vector3d posEva2Rel\(", 0, expect_lines=(515, 530)),
        swm.block(r"^\s*inline Eigen::Vector3d posEva2Rel\(", 1, expect_lines=(525, 540)),
        swm.block(r"^\s*inline double choiceTInit\(const Eigen::Vector3d &pos_eva, double dt\)", expect_lines=(535, 585)),
        swm.block(r"^\s*inline double getSDFAtTimeStamp\(", expect_lines=(738, 760)),
        swm.block(r"^\s*inline Eigen::Vector3d getGradPrelAtTimeStamp\(", expect_lines=(776, 798)),
        swm.block(r"^\s*inline double getSDF_DOTAtTimeStamp\(", expect_lines=(796, 832)),
        swm.block(r"^\s*inline double getSDFofSweptVolume\(const Eigen::Vector3d &pos_eva, double &time_seed_f, "
                  r"Eigen::Vector3d &grad_prel\)\s*$", expect_lines=(842, 868)),
        swm.block(r"^\s*inline double getTrueSDFofSweptVolume\(", expect_lines=(912, 1020)),
        swm.block(r"^\s*inline void gradientDescent\(double momentum", expect_lines=(1245, 1330)),
    ]
    files["sw_methods.inc"] = "".join(methods)

    # --- back_end_optimizer.hpp ------------------------------------------------------------------------------------
    opt = [
        beo.block(r"^\s*static inline void forwardP\(const double\* xi,", expect_lines=(170, 190)),
        beo.block(r"^\s*static inline void forwardT\(const double \*tau,", expect_lines=(212, 232)),
        beo.block(r"^\s*static inline void backwardT\(const Eigen::VectorXd &T,", expect_lines=(226, 246)),
        beo.block(r"^\s*static inline void backwardGradT\(const double\* tau,", expect_lines=(266, 292)),
        beo.block(r"^\s*static inline void backwardGradP\(\s*$", expect_lines=(300, 316)),
        beo.block(r"^\s*static inline bool smoothedL1\(", expect_lines=(314, 342)),
        beo.block(r"^\s*static inline double costFunctionLmbmParallel\(", expect_lines=(340, 410)),
        beo.block(r"^\s*static inline void addSaftyPenaOnSweptVolumeParallelTrueSDF\(", expect_lines=(772, 872)),
        beo.block(r"^\s*bool inline grad_cost_p_sw\(", expect_lines=(1028, 1068)),
    ]
    files["opt_methods.inc"] = "".join(opt)

    # ---- mid end (OriTraj): mid_end.hpp member functions + getOriTraj from mid_end.cpp; utils/flatness.hpp, utils/lbfgs.hpp and
    # utils/minco.hpp are included whole by oracle/ref_mid_shim.cpp ----
    mid = Source(os.path.join(ref, "src/planner_algorithm/include/planner_algorithm/mid_end.hpp"))
    files["mid_methods.inc"] = "".join([
        mid.block(r"^\s*static inline bool smoothedL1\(const double &x,", expect_lines=(60, 92)),
        mid.block(r"^\s*static inline void forwardP\(const Eigen::VectorXd &xi,", expect_lines=(88, 104)),
        mid.block(r"^\s*static inline void backwardP\(const Eigen::Matrix3Xd &P,", expect_lines=(100, 116)),
        mid.block(r"^\s*static inline void forwardT\(const Eigen::VectorXd &tau,", expect_lines=(112, 130)),
        mid.block(r"^\s*static inline void backwardT\(const Eigen::VectorXd &T,", expect_lines=(126, 146)),
        mid.block(r"^\s*static inline void backwardGradT\(const Eigen::VectorXd &tau,", expect_lines=(140, 170)),
        mid.block(r"^\s*static inline void backwardGradP\(const Eigen::VectorXd &xi,", expect_lines=(166, 184)),
        mid.block(r"^\s*bool inline grad_cost_dir\(", expect_lines=(180, 212)),
        mid.block(r"^\s*static inline void addPosePenalty \(", expect_lines=(210, 276)),
        mid.block(r"^\s*static inline double costFunction\(void \*ptr,", expect_lines=(274, 328)),
        mid.block(r"^\s*static inline double costaltitude\(", expect_lines=(370, 394)),
        mid.block(r"^\s*static inline void gradaltitude\(", expect_lines=(390, 420)),
        mid.block(r"^\s*static inline double WC2\(", expect_lines=(414, 436)),
        mid.block(r"^\s*static inline void addTimeIntPenalty\(", expect_lines=(432, 606)),
        mid.block(r"^\s*static inline int earlyExit\(void \*instance,", expect_lines=(598, 626)),
    ])
    midc = Source(os.path.join(ref, "src/planner_algorithm/src/mid_end.cpp"))
    files["mid_getoritraj.inc"] = midc.block(r"^bool OriTraj::getOriTraj\(", expect_lines=(1, 95))

    for name, body in files.items():
        path = os.path.join(out, name)
        with open(path, "w", encoding="utf-8", errors="surrogateescape") as fh:
            fh.write("// GENERATED by oracle/ref_extract.py from the reference tree — verbatim reference source, do not commit.\n")
            fh.write(body)
    print(f"ref_extract: wrote {len(files)} fragments to {out}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_ref", "gen"))
    a = ap.parse_args()
    if not os.path.isdir(a.ref):
        print("ref_extract: no reference tree at", a.ref, file=sys.stderr)
        sys.exit(2)
    extract(a.ref, a.out)
