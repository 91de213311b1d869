#!/usr/bin/env python3
"""Derive the minimal parity-input files under assets/ from the reference checkout.

The GPU box has no /root/reference, so the model/config *inputs* of the hot path travel as
derived fixtures.  This script (run once in the build container) re-serialises only what
the hot path consumes:

  * assets/qm_robot.urdf      <- qm_description/urdf/qudraputed_manipulator/robot.urdf
        links (inertial only), joints (origin/axis/limit); visuals, collisions, gazebo and
        transmission blocks dropped.
  * assets/qm_task.info       <- qm_controllers/config/task.info        (comments stripped,
        the ddp / ipm / rollout blocks are kept for the solver variants)
  * assets/qm_reference.info  <- qm_controllers/config/reference.info   (comments stripped)
  * assets/qm_gait.info       <- qm_controllers/config/gait.info        (comments stripped)
  * assets/qm_wbc_gains.info  <- qm_wbc/cfg/wbcWigeht.cfg               (defaults -> INFO keys)

Usage: python tools/make_assets.py [/root/reference]
"""
import os
import re
import sys
import xml.etree.ElementTree as ET

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def minimise_urdf(src, dst):
    root = ET.parse(src).getroot()
    lines = ['<?xml version="1.0"?>', '<robot name="%s">' % root.get("name", "qm")]
    for el in root:
        if el.tag == "link":
            inert = el.find("inertial")
            if inert is None:
                lines.append('  <link name="%s"/>' % el.get("name"))
                continue
            lines.append('  <link name="%s">' % el.get("name"))
            lines.append("    <inertial>")
            o = inert.find("origin")
            if o is not None:
                lines.append('      <origin xyz="%s" rpy="%s"/>' % (o.get("xyz", "0 0 0"), o.get("rpy", "0 0 0")))
            lines.append('      <mass value="%s"/>' % inert.find("mass").get("value"))
            i = inert.find("inertia")
            lines.append("      <inertia " + " ".join('%s="%s"' % (k, i.get(k)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")) + "/>")
            lines.append("    </inertial>")
            lines.append("  </link>")
        elif el.tag == "joint" and el.get("type") is not None:
            lines.append('  <joint name="%s" type="%s">' % (el.get("name"), el.get("type")))
            lines.append('    <parent link="%s"/>' % el.find("parent").get("link"))
            lines.append('    <child link="%s"/>' % el.find("child").get("link"))
            o = el.find("origin")
            if o is not None:
                lines.append('    <origin xyz="%s" rpy="%s"/>' % (o.get("xyz", "0 0 0"), o.get("rpy", "0 0 0")))
            if el.get("type") != "fixed":
                a = el.find("axis")
                lines.append('    <axis xyz="%s"/>' % (a.get("xyz") if a is not None else "1 0 0"))
                l = el.find("limit")
                if l is not None:
                    lines.append('    <limit lower="%s" upper="%s" effort="%s" velocity="%s"/>' % (
                        l.get("lower", "0"), l.get("upper", "0"), l.get("effort", "0"), l.get("velocity", "0")))
            lines.append("  </joint>")
    lines.append("</robot>")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")


def strip_info(src, dst, drop_blocks=()):
    out, skip_depth, pending_drop = [], 0, False
    for raw in open(src):
        line = re.split(r";|//", raw, 1)[0].rstrip()
        if not line.strip():
            continue
        tok = line.split()
        if skip_depth == 0 and not pending_drop and tok[0] in drop_blocks and len(tok) == 1:
            pending_drop = True
            continue
        if pending_drop:
            if tok[0] == "{":
                pending_drop, skip_depth = False, 1
            continue
        if skip_depth:
            skip_depth += line.count("{") - line.count("}")
            continue
        out.append(re.sub(r"\s+", " ", line.strip()) if "{" not in line and "}" not in line else line.strip())
    # re-indent
    depth, res = 0, []
    for l in out:
        if l.startswith("}"):
            depth -= 1
        res.append("  " * depth + l)
        if l.endswith("{"):
            depth += 1
    with open(dst, "w") as f:
        f.write("\n".join(res) + "\n")


def wbc_gains(src, dst):
    pat = re.compile(r'gen\.add\("(\w+)",\s*double_t,\s*0,\s*"[^"]*",\s*([-\d.eE+]+)')
    rows = [m.groups() for m in map(pat.search, open(src)) if m]
    with open(dst, "w") as f:
        f.write("wbcGains\n{\n")
        for k, v in rows:
            f.write("  %s %s\n" % (k, v))
        f.write("}\n")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    minimise_urdf(os.path.join(REF, "qm_description/urdf/qudraputed_manipulator/robot.urdf"), os.path.join(OUT, "qm_robot.urdf"))
    cfg = os.path.join(REF, "qm_controllers/config")
    strip_info(os.path.join(cfg, "task.info"), os.path.join(OUT, "qm_task.info"), drop_blocks=())   # ddp{} / ipm{} / rollout{} stay: qmb200_mpc_set_solver reads them
    strip_info(os.path.join(cfg, "reference.info"), os.path.join(OUT, "qm_reference.info"))
    strip_info(os.path.join(cfg, "gait.info"), os.path.join(OUT, "qm_gait.info"))
    wbc_gains(os.path.join(REF, "qm_wbc/cfg/wbcWigeht.cfg"), os.path.join(OUT, "qm_wbc_gains.info"))
    print("assets written to", OUT)
