"""MEDIT ``.mesh`` reader (the format of elasticity/data/*.mesh) -- a stand-in for ``meshio.read`` on the one call
the reference makes (elasticity/model.py:77-83: ``mesh.points``, ``mesh.cells_dict['tetra' | 'triangle']``), plus the
mesh set-up of ``ElasticityModel._init_mesh`` (:76-93) for ``fused.ElasticityStepper(mesh=...)``.

Format: whitespace-separated tokens; ``MeshVersionFormatted v``, ``Dimension d``, then blocks ``Vertices n`` (n rows of
d coordinates + a reference tag), ``Triangles n`` (3 indices + tag), ``Tetrahedra n`` (4 indices + tag), ... , ``End``.
Indices are 1-based in the file and 0-based in the returned arrays, like meshio."""
from __future__ import annotations

import numpy as np

# entities of the format: name -> number of vertex indices per row (each row carries one more token, the tag)
_CELLS = {"Edges": ("line", 2), "Triangles": ("triangle", 3), "Quadrilaterals": ("quad", 4), "Tetrahedra": ("tetra", 4),
          "Hexahedra": ("hexahedron", 8)}
# blocks of single indices (no tag) that some writers emit
_INDEX_LISTS = ("Corners", "RequiredVertices", "Ridges", "RequiredEdges")


class Mesh:
    """the two attributes of ``meshio.Mesh`` that the reference reads"""

    def __init__(self, points, cells_dict):
        self.points = points
        self.cells_dict = cells_dict

    def __repr__(self):
        return f"<medit.Mesh {self.points.shape[0]} points, " + ", ".join(f"{v.shape[0]} {k}" for k, v in self.cells_dict.items()) + ">"


def read(path, file_format=None):
    """``meshio.read(path)`` for MEDIT text files"""
    with open(path, "r") as fh:
        text = fh.read()
    # comments run from '#' to the end of the line
    if "#" in text:
        text = "\n".join(line.split("#", 1)[0] for line in text.splitlines())
    tok = text.split()
    if not tok or tok[0] != "MeshVersionFormatted":
        raise ValueError(f"{path}: not a MEDIT text mesh (no MeshVersionFormatted header)")
    pos, dim, points, cells = 2, 3, None, {}
    while pos < len(tok):
        key = tok[pos]
        pos += 1
        if key == "End":
            break
        if key == "Dimension":
            dim = int(tok[pos]); pos += 1
        elif key == "Vertices":
            n = int(tok[pos]); pos += 1
            block = np.array(tok[pos:pos + n * (dim + 1)], dtype=np.float64).reshape(n, dim + 1)
            points = np.ascontiguousarray(block[:, :dim])
            pos += n * (dim + 1)
        elif key in _CELLS:
            name, k = _CELLS[key]
            n = int(tok[pos]); pos += 1
            block = np.array(tok[pos:pos + n * (k + 1)], dtype=np.int64).reshape(n, k + 1)
            cells[name] = np.ascontiguousarray(block[:, :k] - 1)
            pos += n * (k + 1)
        elif key in _INDEX_LISTS:
            n = int(tok[pos]); pos += 1 + n
        else:
            raise ValueError(f"{path}: unknown MEDIT block '{key}'")
    if points is None:
        raise ValueError(f"{path}: no Vertices block")
    return Mesh(points, cells)


def load_normalized(path, dim, device="cpu"):
    """(V, F) as ``ElasticityModel._init_mesh`` prepares them (elasticity/model.py:76-88 with torchgp/normalize.py:22-36):
    vertices centred on the bounding box, scaled so that the farthest vertex sits at distance 1, then doubled;
    F = tetrahedra (dim 3) or triangles (dim 2)."""
    import torch
    mesh = read(path)
    V = torch.tensor(mesh.points, dtype=torch.float32, device=device)
    F = torch.tensor(mesh.cells_dict["tetra" if dim == 3 else "triangle"], device=device)
    center = (V.max(dim=0).values + V.min(dim=0).values) / 2.0
    V = V - center
    V = V * (1.0 / torch.sqrt(torch.max(torch.sum(V ** 2, dim=-1))))
    return V * 2.0, F
