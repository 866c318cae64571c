"""Minimal gmsh ``.msh`` reader/writer for P1 triangle meshes (ASCII formats 2.2 and 4.1).

Stands in for ``dolfinx.io.gmshio.read_from_msh(path, comm, gdim=2)`` of the reference
(setups/setup_cooke2.py:19): returns the mesh container the rest of the host side uses.  Only
what that call needs is implemented: nodes, 3-node triangles (gmsh element type 2) and, like
DOLFINx, only the cells that belong to a physical group when the file defines any (the
reference's meshes write the physical surface "Area", notebooks/create_mesh.ipynb:400).

Node order: gmsh node tags in ascending order become vertices 0..n-1 and the triangles keep the
file's order and vertex order.  DOLFINx additionally reorders cells and dofs for locality when it
distributes a mesh; that permutation is not reproducible offline, so parity with a real FEniCSx
run is established through ``shakti_b200.golden`` (which loads the dofmap DOLFINx actually
used), not through this reader.
"""
import numpy as np

from .fem import Mesh


def _sections(path):
    out, name, buf = {}, None, []
    with open(path, "r") as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            if line.startswith("$End"):
                out[name] = buf
                name, buf = None, []
            elif line.startswith("$"):
                name, buf = line[1:], []
            elif name is not None:
                buf.append(line)
    return out


def read_msh_arrays(path):
    """-> (xy (n,2) float64, cells (m,3) int32, cell_tags (m,) int32)"""
    sec = _sections(path)
    if "MeshFormat" not in sec:
        raise ValueError(f"{path}: not a gmsh .msh file")
    version, ftype = sec["MeshFormat"][0].split()[:2]
    if int(ftype) != 0:
        raise ValueError("binary .msh files are not supported (write ASCII: gmsh -format msh2 / Mesh.Binary=0)")
    major = int(float(version))
    if major == 2:
        n = int(sec["Nodes"][0])
        rows = np.array([l.split() for l in sec["Nodes"][1:1 + n]], dtype=np.float64)
        tags, coords = rows[:, 0].astype(np.int64), rows[:, 1:4]
        tri, ctag = [], []
        for l in sec["Elements"][1:]:
            p = l.split()
            if int(p[1]) != 2:
                continue
            ntags = int(p[2])
            ctag.append(int(p[3]) if ntags > 0 else 0)
            tri.append([int(v) for v in p[3 + ntags:3 + ntags + 3]])
    elif major == 4:
        lines = sec["Nodes"]
        nblocks, n = [int(v) for v in lines[0].split()[:2]]
        tags, coords, i = [], [], 1
        for _ in range(nblocks):
            nb = int(lines[i].split()[3])
            tags += [int(v) for v in lines[i + 1:i + 1 + nb]]
            coords += [[float(v) for v in l.split()[:3]] for l in lines[i + 1 + nb:i + 1 + 2 * nb]]
            i += 1 + 2 * nb
        tags, coords = np.array(tags, dtype=np.int64), np.array(coords, dtype=np.float64)
        lines = sec["Elements"]
        nblocks = int(lines[0].split()[0])
        tri, ctag, i = [], [], 1
        ent2phys = {}
        if "Entities" in sec:
            e = sec["Entities"]
            np_, nc, ns, _ = [int(v) for v in e[0].split()]
            for l in e[1 + np_ + nc:1 + np_ + nc + ns]:
                p = l.split()
                nphys = int(p[7])
                ent2phys[int(p[0])] = int(p[8]) if nphys > 0 else 0
        for _ in range(nblocks):
            dim, ent, etype, nb = [int(v) for v in lines[i].split()]
            if etype == 2:
                for l in lines[i + 1:i + 1 + nb]:
                    tri.append([int(v) for v in l.split()[1:4]])
                    ctag.append(ent2phys.get(ent, 0))
            i += 1 + nb
    else:
        raise ValueError(f"unsupported .msh version {version}")
    if not tri:
        raise ValueError(f"{path}: no 3-node triangles found")
    tri, ctag = np.array(tri, dtype=np.int64), np.array(ctag, dtype=np.int32)
    if (ctag > 0).any():                      # DOLFINx keeps only cells of physical groups
        keep = ctag > 0
        tri, ctag = tri[keep], ctag[keep]
    order = np.argsort(tags)
    lut = np.full(int(tags.max()) + 1, -1, dtype=np.int64)
    lut[tags[order]] = np.arange(tags.size)
    cells = lut[tri]
    if (cells < 0).any():
        raise ValueError(f"{path}: element refers to an unknown node tag")
    used = np.zeros(tags.size, dtype=bool)
    used[cells.ravel()] = True
    xy = coords[order][:, :2]
    if not used.all():                        # drop nodes no kept triangle uses (e.g. geometry points)
        remap = np.cumsum(used) - 1
        cells = remap[cells]
        xy = xy[used]
    return np.ascontiguousarray(xy), np.ascontiguousarray(cells, dtype=np.int32), ctag


def read_from_msh(path, comm=None, gdim=2):
    """``domain, cell_tags, facet_tags = read_from_msh(...)`` like dolfinx.io.gmshio (facet tags: None)."""
    assert gdim == 2
    xy, cells, ctag = read_msh_arrays(path)
    return Mesh(xy, cells, comm), ctag, None


def write_msh(path, xy, cells, physical_tag=1, version=2):
    """ASCII .msh (2.2 or 4.1) with one physical surface; used by the tests."""
    xy, cells = np.asarray(xy, dtype=np.float64), np.asarray(cells, dtype=np.int64)
    with open(path, "w") as f:
        if version == 2:
            f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % xy.shape[0])
            for i, p in enumerate(xy):
                f.write("%d %.17g %.17g 0\n" % (i + 1, p[0], p[1]))
            f.write("$EndNodes\n$Elements\n%d\n" % cells.shape[0])
            for i, c in enumerate(cells):
                f.write("%d 2 2 %d 1 %d %d %d\n" % (i + 1, physical_tag, c[0] + 1, c[1] + 1, c[2] + 1))
            f.write("$EndElements\n")
        else:
            f.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
            f.write("$Entities\n0 0 1 0\n1 %.17g %.17g 0 %.17g %.17g 0 1 %d 0\n$EndEntities\n"
                    % (xy[:, 0].min(), xy[:, 1].min(), xy[:, 0].max(), xy[:, 1].max(), physical_tag))
            f.write("$Nodes\n1 %d 1 %d\n2 1 0 %d\n" % (xy.shape[0], xy.shape[0], xy.shape[0]))
            for i in range(xy.shape[0]):
                f.write("%d\n" % (i + 1))
            for p in xy:
                f.write("%.17g %.17g 0\n" % (p[0], p[1]))
            f.write("$EndNodes\n$Elements\n1 %d 1 %d\n2 1 2 %d\n" % (cells.shape[0], cells.shape[0], cells.shape[0]))
            for i, c in enumerate(cells):
                f.write("%d %d %d %d\n" % (i + 1, c[0] + 1, c[1] + 1, c[2] + 1))
            f.write("$EndElements\n")
