"""How tests/golden/mapping_kats.json was made.

The reference (/root/reference) is a header-only C++ library over OpenVDB/PCL/Eigen, none of which
exist in this image, so no reference code can be executed to *generate* vectors. The only golden
vectors the reference holds for the scan-integration path are the five gtest cases in
/root/reference/tests/mapping.cpp; mapping_kats.json is a hand transcription of those cases
(points, origin, resolution, config, EXPECT_EQ / EXPECT_TRUE / EXPECT_FALSE lines).

This script re-checks the transcription against the reference test source when it is available
(this container), by grepping the literals the JSON depends on. It never runs on the GPU box.
"""
import json, os, re, sys

REF = "/root/reference/tests/mapping.cpp"
here = os.path.dirname(os.path.abspath(__file__))
kats = json.load(open(os.path.join(here, "mapping_kats.json")))
if not os.path.exists(REF):
    print("reference not present; nothing to verify"); sys.exit(0)
src = open(REF).read()
assert len(re.findall(r"^TEST\(Mapping, (\w+)\)", src, re.M)) == len(kats["cases"]) == 5
for name in [c["name"] for c in kats["cases"]]:
    assert f"TEST(Mapping, {name})" in src, name
for lit in ["conf.prob_hit       = 0.9", "conf.prob_miss      = 0.1", "conf.prob_thres_max = 0.51",
            "conf.prob_thres_min = 0.49", "conf.max_range = 0.5", "5 * resolution", "-5 * resolution", "7 * resolution"]:
    assert lit in src, lit
print("mapping_kats.json matches the literals in", REF)
