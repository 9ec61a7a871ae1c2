"""TEST INFRASTRUCTURE ONLY — lifts the loss section of a reference driver's `train()` loop out of the UNMODIFIED file with `ast`:
the statements from `student_hidden = student_outputs['hidden_dict']` to `loss += lagrangian_loss` (Eff_VQA.py:105-176,
Eff_Retrieval.py:101-178, Eff_NLVR.py:100-157, Eff_Captioning.py:99-148), plus the file's own `get_kd_loss` / `get_cor_teacher` /
`soft_cross_entropy`.  The drivers cannot be imported as modules here (ruamel_yaml, datasets, apex at import time), and their loops mix
the loss code with I/O; executing the lifted statements on a pair of model outputs runs the reference's loss mix exactly as written.
Needs /root/reference (tests that use it are skipped without it)."""
import ast
import os
import types

import torch

from . import ref_shim


def lift_train_loss(filename, first="student_hidden", helpers=("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")):
    """Returns run(student_outputs, teacher_outputs, model, global_step, temperature=1.0) -> namespace after the lifted statements
    (`loss` = the step's total, every intermediate term under the driver's own variable name) and the (first, last) source lines."""
    src = open(os.path.join(ref_shim.REF_ROOT, filename)).read()
    tree = ast.parse(src)
    base = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss, "MSELoss": torch.nn.MSELoss}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in helpers]
    exec(compile(ast.Module(body=fns, type_ignores=[]), filename, "exec"), base)
    train = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train")

    def assigns(s, name):
        return isinstance(s, ast.Assign) and getattr(s.targets[0], "id", None) == name
    loop = next(n for n in ast.walk(train) if isinstance(n, ast.For) and any(assigns(s, first) for s in n.body))
    i0 = next(i for i, s in enumerate(loop.body) if assigns(s, first))
    i1 = next(i for i, s in enumerate(loop.body) if isinstance(s, ast.AugAssign) and getattr(s.target, "id", None) == "loss"
              and getattr(s.value, "id", None) == "lagrangian_loss")
    stmts = loop.body[i0:i1 + 1]
    code = compile(ast.Module(body=stmts, type_ignores=[]), filename, "exec")
    lines = (stmts[0].lineno, stmts[-1].end_lineno)

    def run(student_outputs, teacher_outputs, model, global_step, temperature=1.0, device="cpu", **extra):
        ns = dict(base)
        ns.update(student_outputs=student_outputs, teacher_outputs=teacher_outputs, model=types.SimpleNamespace(module=model),
                  global_step=global_step, device=device, args=types.SimpleNamespace(temperature=temperature),
                  optimizer=types.SimpleNamespace(zero_grad=lambda: None),     # Eff_NLVR.py:151 / Eff_Captioning.py:143 zero it mid-way
                  **extra)
        exec(code, ns)
        return ns
    return run, lines
