"""Compiled programs: what `model_dict[mode][label]` holds instead of an SDD object.

Reference contract (models/vael.py:98-135, :245-266): `sdd.evaluate(semiring=GraphSemiring)`
returns an ordered mapping {Term -> Tensor[B]} containing Term('digits')(Constant(j),
Constant(k)) for every world and, in query mode, the addition query as the LAST key.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Hashable, Optional

import torch

from .compiler import CompiledFamily, ProgramView
from .ops import DeviceCircuit, circuit_forward


class EvaluationResult(OrderedDict):
    """{Term: Tensor[B]} plus the dense tensors the columns are views of."""
    worlds: Optional[torch.Tensor] = None   # [B,W]
    query: Optional[torch.Tensor] = None    # [B,1]


class CompiledProgramSet:
    """One compiled circuit shared by a family of programs; uploaded lazily, once per GPU."""

    def __init__(self, family: CompiledFamily, device=None):
        self.family = family
        self._dc: Dict[int, DeviceCircuit] = {}
        self._views: Dict[Hashable, "CompiledProgram"] = {}
        if device is not None and torch.device(device).type == "cuda":
            self.device_circuit(torch.device(device))

    @property
    def circuit(self):
        return self.family.circuit

    def __deepcopy__(self, memo):  # compiled programs are immutable: copies of a model share them
        return self

    def __getstate__(self):  # device handles do not travel; they are re-created on first use
        return {"family": self.family}

    def __setstate__(self, state):
        self.family, self._dc, self._views = state["family"], {}, {}

    def device_circuit(self, device=None) -> DeviceCircuit:
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("vael_b200: compiled programs are evaluated on a CUDA device (no CPU fallback)")
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if idx not in self._dc:
            self._dc[idx] = DeviceCircuit(self.family.circuit, torch.device("cuda", idx))
        return self._dc[idx]

    def view(self, key) -> "CompiledProgram":
        if key not in self._views:
            self._views[key] = CompiledProgram(self, self.family.views[key])
        return self._views[key]


class CompiledProgram:
    def __init__(self, pset: CompiledProgramSet, view: ProgramView):
        self.pset, self.view = pset, view

    def __deepcopy__(self, memo):
        return self

    @property
    def program(self) -> str:
        return self.view.text

    @property
    def mode(self) -> str:
        return self.view.mode

    def evaluate(self, semiring=None, **_ignored) -> EvaluationResult:
        """Weighted model counts of every query atom, normalised by Z (evidence programs:
        by the evidence's own count), for the weights bound to `semiring`."""
        if semiring is None:
            raise ValueError("evaluate needs the GraphSemiring carrying the fact weights")
        fam = self.pset.family
        if hasattr(semiring, "fact_matrix"):
            facts = semiring.fact_matrix(fam.fact_terms)
        else:  # a foreign semiring (the reference's own GraphSemiring): read the bound leaf weights once
            from .graph_semiring import GraphSemiring
            facts = GraphSemiring.fact_matrix(semiring, fam.fact_terms)
        dc = self.pset.device_circuit(facts.device)
        evid = self.view.mode == "evidence"
        q = self.view.query_index if self.view.query_index >= 0 else None
        if evid:
            with torch.no_grad():
                worlds, query = circuit_forward(dc, facts, q, evidence=True)
        else:
            worlds, query = circuit_forward(dc, facts, q, normalize=True)
        res = EvaluationResult()
        for i, atom in enumerate(fam.world_terms):
            res[atom] = worlds[:, i]
        for atom in self.view.result_atoms[len(fam.world_terms):]:
            res[atom] = query[:, 0]
        res.worlds, res.query = worlds, query
        return res
