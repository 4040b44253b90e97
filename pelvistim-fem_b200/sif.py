"""Elmer Solver Input File (``case.sif``) model: build, serialise, parse.

The SIF is the problem-definition half of the drop-in boundary (the mesh
directory is the other half).  The reference writes it from string templates
(``step01_box/setup_case.py:139-211``, ``step02_electrodes/run_sweep.py:197-272``,
``step03_ankle_layers/run_layered_sweep.py:459-633``,
``step04_pressure/run_pressure_sweep.py:297-432``); here a case is a list of
``Section`` objects that is serialised to the same text, byte for byte (checked
against the reference's committed ``case.sif`` files in ``tests/golden``), and
``parse_sif``/``problem_from_sif`` read the subset of keywords the engine acts on
(SURVEY.md section 8b).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field


@dataclass
class Section:
    kind: str                     # "Header", "Body", "Material", ...
    index: int | None = None
    lines: list = field(default_factory=list)   # str (already "Key = value") or None (blank line)
    lead: str | None = None       # comment line printed directly above the section

    def add(self, key, value=None, comment=None, gap=2):
        if value is None:
            s = key
        else:
            s = f"{key} = {value}"
        if comment:
            s += " " * gap + f"! {comment}"
        self.lines.append(s)
        return self

    def blank(self):
        self.lines.append(None)
        return self


def serialize(sections):
    out = []
    for n, s in enumerate(sections):
        if n:
            out.append("")
        if s.lead:
            out.append(s.lead)
        out.append(s.kind if s.index is None else f"{s.kind} {s.index}")
        for ln in s.lines:
            out.append("" if ln is None else "  " + ln)
        out.append("End")
    return "\n".join(out) + "\n"


# -- common blocks -----------------------------------------------------------
def _header():
    return (Section("Header").add("CHECK KEYWORDS Warn").add('Mesh DB "." "elmer_mesh"')
            .add('Include Path ""').add('Results Directory "results"'))


def _simulation(level):
    return (Section("Simulation").add("Max Output Level", level).add("Coordinate System", "Cartesian 3D")
            .add("Coordinate Mapping(3)", "1 2 3").add("Simulation Type", "Steady State")
            .add("Steady State Max Iterations", 1).add("Output Intervals", 1))


def _constants():
    return Section("Constants").add("Permittivity of Vacuum", "8.8542e-12")


def _equation():
    return Section("Equation", 1).add("Name", '"Conduction"').add("Active Solvers(2)", "1 2")


def _solver_current(tol="1.0e-8", method="UMFPACK", spaced=False):
    s = Section("Solver", 1)
    s.add("Equation", '"Static Current Conduction"').add("Procedure", '"StatCurrentSolve" "StatCurrentSolver"')
    s.add("Variable", '"Potential"').add("Variable DOFs", 1)
    if spaced:
        s.blank()
    s.add("Calculate Volume Current", "True")
    if spaced:
        s.blank()
    s.add("Linear System Solver", "Direct").add("Linear System Direct Method", method)
    if spaced:
        s.blank()
    s.add("Steady State Convergence Tolerance", tol)
    return s


def _solver_output():
    return (Section("Solver", 2).add("Equation", '"ResultOutput"')
            .add("Procedure", '"ResultOutputSolve" "ResultOutputSolver"').add("Output File Name", '"case"')
            .add("Output Format", "VTU").add("VTU Format", "Logical True").add("Save Geometry IDs", "Logical True"))


def format_target(ids):
    """``Target Boundaries = k`` or ``Target Boundaries(N) = k1 k2 ...``
    (``step01_box/setup_case.py:133-137``)."""
    ids = sorted(set(int(i) for i in ids))
    if len(ids) == 1:
        return f"Target Boundaries = {ids[0]}"
    return f"Target Boundaries({len(ids)}) = " + " ".join(str(i) for i in ids)


def _dirichlet_bc(index, name, target_line, value):
    return Section("Boundary Condition", index).add("Name", f'"{name}"').add(target_line).add("Potential", value)


# -- step01 -------------------------------------------------------------------
def box_case(top_ids, bot_ids, sigma=0.2, v_top=1.0, v_bot=0.0):
    """Homogeneous box, Dirichlet top/bottom: the case ``setup_case.py`` writes."""
    body = (Section("Body", 1).add("Target Bodies(1)", 1).add("Name", '"body"').add("Equation", 1).add("Material", 1))
    mat = Section("Material", 1).add("Name", '"Tissue"').add("Electric Conductivity", sigma)
    return [_header(), _simulation(5), _constants(), body, _equation(), _solver_current(spaced=True),
            _solver_output(), mat,
            _dirichlet_bc(1, "top", format_target(top_ids), v_top),
            _dirichlet_bc(2, "bottom", format_target(bot_ids), v_bot)]


# -- step02 -------------------------------------------------------------------
def electrode_case(e1_id, e2_id, sigma=0.2, v_active=1.0, v_return=0.0):
    """One tissue body, two Dirichlet electrode patches (``run_sweep.py:197-272``)."""
    body = (Section("Body", 1).add("Target Bodies(1)", 1).add("Name", '"tissue"').add("Equation", 1).add("Material", 1))
    mat = Section("Material", 1).add("Name", '"Tissue"').add("Electric Conductivity", sigma)
    return [_header(), _simulation(3), _constants(), body, _equation(), _solver_current(), _solver_output(), mat,
            _dirichlet_bc(1, "active", f"Target Boundaries = {e1_id}", v_active),
            _dirichlet_bc(2, "return", f"Target Boundaries = {e2_id}", v_return)]


# -- step03 / step04 ----------------------------------------------------------
def electrode_area_analytic(elec_r, shape):
    return math.pi * elec_r ** 2 if shape == "circle" else (2 * elec_r) ** 2


def layered_case(e1_id, e2_id, sigma_muscle, sigma_fat, sigma_skin, sigma_contact, contact=True,
                 c1_body=4, c2_body=5, mode="current", injected_current_mA=5.0, elec_r=0.010,
                 shape="circle", elec_area_mesh=None, tol=1e-8, lin_solver="UMFPACK", dialect="step03",
                 sigma_bone=None, bone_body=6):
    """Layered-slab case (``run_layered_sweep.py:507-633``; step04 variant
    ``run_pressure_sweep.py:297-432`` selected with ``dialect='step04'``).

    Returns ``(sections, jn_used, warning)``; ``jn_used`` = I / A (A = mesh area of the
    active electrode when given, else the analytic footprint area), ``None`` in
    voltage mode.  ``sigma_bone``: adds a bone body (mesh body ``bone_body``) with its own material after the
    reference's bodies / materials (not in the reference's SIFs, which are reproduced byte for byte without it)."""
    secs = [_header(), _simulation(3), _constants(), _equation(),
            _solver_current(tol=tol, method=lin_solver), _solver_output()]
    for i, (name, mat) in enumerate((("muscle", 1), ("fat", 2), ("skin", 3)), start=1):
        secs.append(Section("Body", i).add("Name", f'"{name}"').add("Target Bodies(1)", i)
                    .add("Equation", 1).add("Material", mat))
    if contact:
        secs.append(Section("Body", 4).add("Name", '"contact_active"').add("Target Bodies(1)", c1_body)
                    .add("Equation", 1).add("Material", 4))
        secs.append(Section("Body", 5).add("Name", '"contact_return"').add("Target Bodies(1)", c2_body)
                    .add("Equation", 1).add("Material", 4))
    if sigma_bone is not None:
        secs.append(Section("Body", 6 if contact else 4).add("Name", '"bone"').add("Target Bodies(1)", bone_body)
                    .add("Equation", 1).add("Material", 5 if contact else 4))
    m1 = Section("Material", 1).add("Name", '"muscle"').add("Electric Conductivity", sigma_muscle)
    if dialect == "step03":
        m1.lead = "! PLACEHOLDER conductivities — replace with measured values"
    secs += [m1, Section("Material", 2).add("Name", '"fat"').add("Electric Conductivity", sigma_fat),
             Section("Material", 3).add("Name", '"skin"').add("Electric Conductivity", sigma_skin)]
    if contact:
        note = ("effective contact conductivity — PLACEHOLDER" if dialect == "step03"
                else "pressure-dependent — PLACEHOLDER")
        secs.append(Section("Material", 4).add("Name", '"contact"')
                    .add("Electric Conductivity", sigma_contact, comment=note, gap=3))
    if sigma_bone is not None:
        secs.append(Section("Material", 5 if contact else 4).add("Name", '"bone"').add("Electric Conductivity", sigma_bone))
    jn_used = None
    warning = None
    bc1 = Section("Boundary Condition", 1).add("Name", '"active_electrode"').add(f"Target Boundaries = {e1_id}")
    if mode == "voltage":
        bc1.add("Potential", "1.0")
    else:
        I_A = injected_current_mA * 1e-3
        area_an = electrode_area_analytic(elec_r, shape)
        if elec_area_mesh is not None and elec_area_mesh > 0:
            area = elec_area_mesh
            rel = abs(area - area_an) / area_an
            if rel > 0.10:
                warning = (f"mesh electrode area ({area*1e4:.4f} cm²) differs {rel:.1%} "
                           f"from analytic ({area_an*1e4:.4f} cm²)")
        else:
            area = area_an
        jn_used = I_A / area
        bc1.add("Current Density", f"{jn_used:.6e}", comment=f"I={I_A*1e3:.1f}mA, A_mesh={area*1e4:.4f}cm²",
                gap=2 if dialect == "step03" else 3)
    secs.append(bc1)
    secs.append(_dirichlet_bc(2, "return_electrode", f"Target Boundaries = {e2_id}", "0.0"))
    return secs, jn_used, warning


# -- parser -------------------------------------------------------------------
_SEC_RE = re.compile(r"^(Header|Simulation|Constants|Body Force|Body|Equation|Solver|Material|"
                     r"Boundary Condition|Initial Condition)\s*(\d+)?\s*$", re.IGNORECASE)


def parse_sif(text):
    """Parse SIF text into ``[(kind, index, {key_lower: value_str})]``.

    Comments (``! ...``) are stripped; ``Key(N) = v1 v2`` keeps ``key`` without the
    size suffix; keyword-only lines (``Mesh DB "." "elmer_mesh"``) are split on the
    first quote/space into key and value."""
    sections = []
    cur = None
    for raw in text.splitlines():
        line = raw.split("!", 1)[0].strip()
        if not line:
            continue
        if cur is None:
            m = _SEC_RE.match(line)
            if not m:
                raise ValueError(f"SIF: unexpected line outside a section: {raw!r}")
            cur = (m.group(1).title(), int(m.group(2)) if m.group(2) else None, {})
            continue
        if line.lower() == "end":
            sections.append(cur)
            cur = None
            continue
        if "=" in line:
            k, v = line.split("=", 1)
            k = re.sub(r"\(\s*\d+\s*\)\s*$", "", k.strip()).strip().lower()
            cur[2][k] = v.strip()
        else:
            m = re.match(r'^([A-Za-z ]+?)\s+(".*|\S+)$', line)
            if m:
                cur[2][m.group(1).strip().lower()] = m.group(2).strip()
            else:
                cur[2][line.lower()] = ""
    if cur is not None:
        raise ValueError("SIF: missing End")
    return sections


def _num(s):
    s = s.strip()
    for pre in ("real", "integer"):
        if s.lower().startswith(pre + " "):
            s = s[len(pre):].strip()
    return float(s.replace("d", "e").replace("D", "e")) if re.search(r"\d[dD][+-]?\d", s) else float(s)


@dataclass
class Problem:
    mesh_db: str
    results_dir: str
    output_name: str
    sigma_by_body: dict            # body id (mesh "Target Bodies") -> conductivity
    dirichlet: list                # [(boundary id, value)]
    neumann: list                  # [(boundary id, current density)]
    calc_current: bool


def problem_from_sif(text) -> Problem:
    """Reduce a parsed SIF to what the engine needs."""
    secs = parse_sif(text)
    mesh_db, results_dir, out_name = "elmer_mesh", "results", "case"
    materials, bodies = {}, []
    dirichlet, neumann = [], []
    calc_current = False
    for kind, idx, kv in secs:
        if kind == "Header":
            if "mesh db" in kv:
                parts = re.findall(r'"([^"]*)"', kv["mesh db"])
                if len(parts) >= 2:
                    mesh_db = parts[1] if parts[0] in (".", "") else f"{parts[0]}/{parts[1]}"
            if "results directory" in kv:
                results_dir = kv["results directory"].strip('"')
        elif kind == "Material":
            if "electric conductivity" in kv:
                materials[idx] = _num(kv["electric conductivity"])
        elif kind == "Body":
            tb = kv.get("target bodies")
            targets = [int(t) for t in tb.split()] if tb else [idx]
            bodies.append((targets, int(kv["material"]) if "material" in kv else None))
        elif kind == "Boundary Condition":
            tb = kv.get("target boundaries")
            if tb is None:
                continue
            ids = [int(t) for t in tb.split()]
            if "potential" in kv:
                dirichlet += [(i, _num(kv["potential"])) for i in ids]
            if "current density" in kv:
                neumann += [(i, _num(kv["current density"])) for i in ids]
        elif kind == "Solver":
            if kv.get("calculate volume current", "").lower().endswith("true"):
                calc_current = True
            if "output file name" in kv:
                out_name = kv["output file name"].strip('"')
    sigma_by_body = {}
    for targets, mat in bodies:
        if mat is None or mat not in materials:
            raise ValueError(f"SIF: body {targets} has no material with an Electric Conductivity")
        for t in targets:
            sigma_by_body[t] = materials[mat]
    if not sigma_by_body:
        raise ValueError("SIF: no Body/Material conductivity found")
    return Problem(mesh_db, results_dir, out_name, sigma_by_body, dirichlet, neumann, calc_current)
