"""Attribute-bag stand-in for `acados_template.AcadosModel` used while lowering the
reference model to CUDA (codegen/gen_model.py).  The solver-facing classes live in
b200mpc/acados_compat.py."""


class AcadosModel:
    def __init__(self):
        self.name = None
        self.x = self.xdot = self.u = self.z = self.p = None
        self.f_impl_expr = self.f_expl_expr = None
        self.con_h_expr = self.con_h_expr_e = None
