"""Golden fixture for the host-side Parameters mirror, derived from the reference tree (run HERE, where
/root/reference exists; the fixture travels, the reference does not):

    python tests/golden/make_params_golden.py

* defaults:  every `p*` member initialiser of cppsrc/core/include/CoverageControl/parameters.h:49-175
* toml:      the leaf values of params/coverage_control_params.toml (the TOML eval.toml and
             data_params.toml point at), keyed by the `p*` attribute parameters.cpp:36-261 loads them into
* toml_keys: every leaf key of that file (the reference's python/tests/test_parameters.py checks that every
             `p*` attribute, minus the p, is one of them)
"""
import json
import os
import re
import tomllib

REF = os.environ.get("CC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def header_defaults():
    src = open(os.path.join(REF, "cppsrc/core/include/CoverageControl/parameters.h")).read()
    out = {}
    for typ, name, val in re.findall(r"^\s*(int|double|bool|float)\s+(p[A-Za-z]+)\s*=\s*([^;]+);", src, flags=re.M):
        val = val.strip()
        if typ == "bool":
            out[name] = val == "true"
        else:  # literals, or an earlier member (pRobotMapSize = pWorldMapSize)
            v = eval(val, {}, dict(out))
            out[name] = int(v) if typ == "int" else float(v)
    return out


def leaves(node, prefix=()):
    for k, v in node.items():
        if isinstance(v, dict):
            yield from leaves(v, prefix + (k,))
        else:
            yield prefix + (k,), v


def main():
    with open(os.path.join(REF, "params/coverage_control_params.toml"), "rb") as f:
        cfg = tomllib.load(f)
    leaf = dict(leaves(cfg))
    by_key = {}
    for path, v in leaf.items():
        by_key.setdefault(path[-1], []).append(v)
    defaults = header_defaults()
    toml = {}
    for attr in defaults:
        vals = by_key.get(attr[1:])
        if vals and len(vals) == 1:
            toml[attr] = vals[0]
    out = {"defaults": defaults, "toml": toml, "toml_keys": sorted({p[-1] for p in leaf}),
           "source": "parameters.h:49-175, params/coverage_control_params.toml of the reference tree"}
    with open(os.path.join(HERE, "params_golden.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"{len(defaults)} defaults, {len(toml)} toml values, {len(out['toml_keys'])} toml keys")


if __name__ == "__main__":
    main()
