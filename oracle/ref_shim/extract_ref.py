#!/usr/bin/env python
"""Build-time extraction of the reference's own function bodies (TEST INFRASTRUCTURE, see oracle/ref_shim/README.md).

Reads /root/reference/MonoSLAM/SLAM.cpp, finds every function named in FUNCTIONS by its definition line
(`<return type> CSLAM::<name> (`) and copies the text from that line to the matching closing brace VERBATIM into
oracle/_ref/slam_extract.cpp (git-ignored build directory; nothing of the reference is committed).  The extracted text is
compiled against oracle/ref_shim/ref_shim.h, a small stand-in for the cv::Mat / GSL / MFC surface those bodies use.
Prints a manifest (function, first line, last line, sha1 of the text) that the tests compare with the line ranges cited in
SURVEY.md / DESIGN.md.
"""
import hashlib
import os
import re
import sys

FUNCTIONS = [
    "CSLAM",                       # constructor: the FLAG_* / EPSILON / CHI2INV_TABLE constants, SLAM.cpp:21-56
    "initializeParameters",        # camera, noise and prior parameters, :158-343
    "getTransferMatrix", "calculateSampleParameter", "expandMatrix", "generateSigmaPoints",
    "passSigmaThroughMapingFunction", "QrAndCholeskyForInitilization", "getPermutationMatrix",
    "predictMotion", "passSigmaThroughMotionFunction", "QrAndCholeskyForMotion",
    "predictMeasurement", "passSigmaThroughMesaurementFunction", "QrAndCholeskyForMeasurement",
    "calculateOneFeatureCovariance", "calculateOneFeatureCrossCovariance", "KalmanUpdate",
    "GSLCholeskyUpdate", "CholeskyDecompositionWithPivoting", "modifiedCholeskyDecomposition", "GSLQrDecomposition",
    "deleteOneFeature",
    "distortOnePointRW", "undistortOnePointRW", "coordinatesState2World", "coordinatesWorld2Camera",
    "coordinatesCamera2Image", "coordinatesImage2Camera", "coordinatesCamera2World", "coordinatesWorld2State",
    "dataTypeCVMat2GSLMat", "dataTypeGSLMat2CVMat",
]


def extract(src_path):
    text = open(src_path, encoding="utf-8-sig", errors="replace").read()
    lines = text.split("\n")
    out, manifest = [], []
    for name in FUNCTIONS:
        pat = re.compile(r"^[A-Za-z_][\w\s\*&:<>]*\bCSLAM::" + re.escape(name) + r"\s*\(") if name != "CSLAM" else \
            re.compile(r"^CSLAM::CSLAM\s*\(")
        start = next((i for i, l in enumerate(lines) if pat.match(l)), None)
        if start is None:
            raise SystemExit(f"extract_ref: definition of CSLAM::{name} not found in {src_path}")
        depth, end, seen = 0, None, False
        for i in range(start, len(lines)):
            code = re.sub(r"//.*", "", lines[i])
            code = re.sub(r'"(\\.|[^"\\])*"', '""', code)
            code = re.sub(r"'(\\.|[^'\\])*'", "''", code)
            for ch in code:
                if ch == "{":
                    depth += 1
                    seen = True
                elif ch == "}":
                    depth -= 1
            if seen and depth == 0:
                end = i
                break
        if end is None:
            raise SystemExit(f"extract_ref: unbalanced braces in CSLAM::{name}")
        body = "\n".join(lines[start:end + 1])
        if "/*" in re.sub(r"//.*", "", body) and "*/" not in body:
            raise SystemExit(f"extract_ref: open block comment in CSLAM::{name}")
        out.append(f"// ---- SLAM.cpp:{start + 1}-{end + 1} (verbatim) ----\n{body}\n")
        manifest.append((name, start + 1, end + 1, hashlib.sha1(body.encode()).hexdigest()[:12]))
    return "\n".join(out), manifest


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/MonoSLAM"
    dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "_ref")
    os.makedirs(dst, exist_ok=True)
    body, manifest = extract(os.path.join(ref, "SLAM.cpp"))
    with open(os.path.join(dst, "slam_extract.cpp"), "w", encoding="utf-8") as f:
        f.write("// GENERATED at build time by oracle/ref_shim/extract_ref.py from the reference's SLAM.cpp -- do not commit.\n")
        f.write('#include "ref_shim.h"\n#include "SLAM.h"   // the reference\'s own header, copied verbatim beside this file\n\n')
        f.write(body)
    # the reference's own header, verbatim, next to the extracted bodies (its `#include "CvImage.h"` then resolves to
    # the stand-in under ref_shim/ instead of the MFC one beside the original)
    with open(os.path.join(ref, "SLAM.h"), "rb") as f:
        hdr = f.read()
    with open(os.path.join(dst, "SLAM.h"), "wb") as f:
        f.write(hdr)
    with open(os.path.join(dst, "manifest.txt"), "w") as f:
        for name, a, b, h in manifest:
            f.write(f"{name} {a} {b} {h}\n")
    for name, a, b, h in manifest:
        print(f"  CSLAM::{name:36s} SLAM.cpp:{a}-{b}  sha1 {h}")


if __name__ == "__main__":
    main()
