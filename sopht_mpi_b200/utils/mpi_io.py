"""Checkpoint / restart of the Eulerian and Lagrangian state.

Same registration API as the reference's ``MPIIO`` (``sopht_mpi/utils/mpi_io.py:7-554``:
``define_eulerian_grid``, ``add_as_eulerian_fields_for_io``, ``add_as_lagrangian_fields_for_io``,
``save(file_name, time)``, ``load(file_name) -> time``), so the save / restart recipe of
``examples/3d_examples/FlowPastFreelyRotatingRodCase/flow_past_freely_rotating_rod_case.py:196-247`` reads the
same.  What differs is the container: ``h5py`` is not part of this image, so a checkpoint is ONE numpy
``.npz`` archive written by rank 0 (global interior of every Eulerian field, gathered from the slabs;
Lagrangian fields from their master rank) instead of a parallel HDF5 file, and no XDMF companion is
generated (``generate_xdmf_*`` are no-ops).  Fields may be numpy arrays or device fields: saving reads
the interior back from the GPU, loading writes it through the field's ``__setitem__``.
"""
import numpy as np

from .logger import logger


class MPIIO:
    def __init__(self, mpi_construct, real_dtype=np.float64):
        self.mpi_construct = mpi_construct
        self.dim = mpi_construct.grid_dim
        assert self.dim == 2 or self.dim == 3, "Invalid dimension (only 2D and 3D)"
        self.real_dtype = real_dtype
        self.precision = 8 if real_dtype is np.float32 else 16
        self.eulerian_grid_defined = False
        self.eulerian_fields = {}
        self.eulerian_fields_type = {}
        self.lagrangian_fields = {}
        self.lagrangian_fields_type = {}
        self.lagrangian_grids = {}
        self.lagrangian_fields_with_grid_name = {}
        self.lagrangian_grid_count = 0
        self.lagrangian_grid_connection = {}
        self.lagrangian_grid_master_rank = {}
        self.lagrangian_grid_num_node = {}

    # ------------------------------------------------------------------ registration
    def define_eulerian_grid(self, origin, dx, grid_size, ghost_size):
        """origin, dx, grid_size: (dim,) arrays in z-y-x order (reference :60-120)"""
        assert isinstance(origin, np.ndarray)
        assert isinstance(dx, np.ndarray)
        assert isinstance(grid_size, np.ndarray)
        if ghost_size < 0 and not isinstance(ghost_size, int):
            raise ValueError(f"Ghost size {ghost_size} needs to be an integer >= 0for eulerian field IO.")
        self.eulerian_origin, self.eulerian_dx, self.eulerian_grid_size = origin, dx, grid_size
        self.eulerian_grid_defined = True
        mc = self.mpi_construct
        start = np.asarray(mc.grid.coords) * np.asarray(mc.local_grid_size)
        end = (np.asarray(mc.grid.coords) + 1) * np.asarray(mc.local_grid_size)
        self.local_eulerian_index = (...,) + tuple(slice(int(a), int(b)) for a, b in zip(start, end))
        self.ghost_size = ghost_size
        self.local_eulerian_grid_size = mc.local_grid_size
        self.local_eulerian_grid_size_with_ghost = self.local_eulerian_grid_size + 2 * ghost_size
        self.eulerian_field_inner_index = (... if ghost_size == 0
                                           else (slice(ghost_size, -ghost_size),) * self.dim)

    def add_as_eulerian_fields_for_io(self, **fields_for_io):
        assert self.eulerian_grid_defined, "Eulerian mesh is not defined!"
        for name, field in fields_for_io.items():
            assert np.issubdtype(field.dtype, self.real_dtype), (
                f"{name} dtype ({field.dtype}) incompatible with IO dtype ({self.real_dtype})")
            shape = tuple(int(s) for s in field.shape)
            local = tuple(int(s) for s in self.local_eulerian_grid_size_with_ghost)
            if shape == local:
                self.eulerian_fields_type[name] = "Scalar"
            elif shape == (self.dim,) + local:
                self.eulerian_fields_type[name] = "Vector"
            else:
                raise ValueError("Unable to identify eulerian field type "
                                 f"(scalar / vector) based on field dimension {field.shape}")
            self.eulerian_fields[name] = field

    def add_as_lagrangian_fields_for_io(self, lagrangian_grid_master_rank, lagrangian_grid,
                                        lagrangian_grid_name=None, lagrangian_grid_connect=False,
                                        **fields_for_io):
        assert len(lagrangian_grid.shape) == 2, "lagrangian grid has to be a 2D (dim, N) array."
        assert lagrangian_grid.shape[0] == self.dim, "Invalid lagrangian grid dimension (only 2D and 3D)"
        if lagrangian_grid_name is None:
            lagrangian_grid_name = f"Lagrangian_grid_{self.lagrangian_grid_count}"
            self.lagrangian_grid_count += 1
        grid = self.mpi_construct.grid
        n = grid.bcast(lagrangian_grid.shape[1], root=lagrangian_grid_master_rank)
        self.lagrangian_grid_num_node[lagrangian_grid_name] = n
        self.lagrangian_grid_master_rank[lagrangian_grid_name] = lagrangian_grid_master_rank
        if lagrangian_grid_connect:
            self.lagrangian_grid_connection[lagrangian_grid_name] = np.arange(n, dtype=np.int64)
        assert np.issubdtype(lagrangian_grid.dtype, self.real_dtype), (
            f"{lagrangian_grid_name} dtype ({lagrangian_grid.dtype}) incompatible with IO dtype ({self.real_dtype})")
        self.lagrangian_grids[lagrangian_grid_name] = lagrangian_grid
        self.lagrangian_fields_with_grid_name[lagrangian_grid_name] = []
        for name, field in fields_for_io.items():
            assert np.issubdtype(field.dtype, self.real_dtype), (
                f"{name} dtype ({field.dtype}) incompatible with IO dtype ({self.real_dtype})")
            self.lagrangian_fields[name] = field
            self.lagrangian_fields_with_grid_name[lagrangian_grid_name].append(name)
            kind = None
            if self.mpi_construct.rank == lagrangian_grid_master_rank:
                if field.shape[0] == lagrangian_grid.shape[1]:
                    kind = "Scalar"
                elif field.shape == lagrangian_grid.shape:
                    kind = "Vector"
                else:
                    raise ValueError("Unable to identify lagrangian field type "
                                     f"(scalar / vector) based on field dimension {field.shape}")
            self.lagrangian_fields_type[name] = grid.bcast(kind, root=lagrangian_grid_master_rank)

    # ------------------------------------------------------------------ save / load
    def _gather_eulerian(self, field):
        """global interior of a local (padded) field on rank 0, None elsewhere"""
        mc = self.mpi_construct
        local = np.ascontiguousarray(np.asarray(field[(...,) + tuple(self.eulerian_field_inner_index)]
                                                if self.ghost_size else field))
        if mc.size == 1:
            return local
        blocks = mc.grid.allgather(local)
        if mc.rank != 0:
            return None
        lead = local.shape[:local.ndim - self.dim]
        out = np.empty(lead + tuple(int(v) for v in self.eulerian_grid_size), dtype=local.dtype)
        for r, block in enumerate(blocks):
            coords = mc.grid.Get_coords(r)
            sl = tuple(slice(int(c * n), int((c + 1) * n)) for c, n in zip(coords, mc.local_grid_size))
            out[(...,) + sl] = block
        return out

    def save(self, h5_file_name, time=0.0):
        self._save(h5_file_name, time)

    def _save(self, file_name, time=0.0):
        mc = self.mpi_construct
        payload = {"time": np.asarray(time, dtype=np.float64)}
        if self.eulerian_grid_defined:
            payload["Eulerian/Parameters/origin"] = np.asarray(self.eulerian_origin)
            payload["Eulerian/Parameters/dx"] = np.asarray(self.eulerian_dx)
            payload["Eulerian/Parameters/grid_size"] = np.asarray(self.eulerian_grid_size)
            for name, field in self.eulerian_fields.items():
                data = self._gather_eulerian(field)
                if data is not None:
                    payload[f"Eulerian/{self.eulerian_fields_type[name]}/{name}"] = data
        for grid_name, names in self.lagrangian_fields_with_grid_name.items():
            master = self.lagrangian_grid_master_rank[grid_name]
            items = None
            if mc.rank == master:
                items = {f"Lagrangian/{grid_name}/Grid": np.array(self.lagrangian_grids[grid_name])}
                for name in names:
                    items[f"Lagrangian/{grid_name}/{self.lagrangian_fields_type[name]}/{name}"] = np.array(
                        self.lagrangian_fields[name])
            if master != 0:
                items = mc.grid.bcast(items, root=master)
            if mc.rank == 0 and items:
                payload.update(items)
        if mc.rank == 0:
            with open(file_name, "wb") as fh:  # (keep the caller's file name, whatever its extension)
                np.savez(fh, **payload)
        mc.grid.Barrier()

    def load(self, h5_file_name):
        """fill the registered fields from a checkpoint; returns its time"""
        mc = self.mpi_construct
        with np.load(h5_file_name) as data:
            time = float(data["time"])
            if self.eulerian_grid_defined:
                for key, mine in (("origin", self.eulerian_origin), ("dx", self.eulerian_dx),
                                  ("grid_size", self.eulerian_grid_size)):
                    if not np.allclose(data[f"Eulerian/Parameters/{key}"], mine):
                        logger.error(f"checkpoint {key} differs from the Eulerian grid defined for IO")
                        raise ValueError("Inconsistent Eulerian grid between checkpoint and simulator")
                inner = ((...,) + tuple(self.eulerian_field_inner_index)) if self.ghost_size else ...
                for name, field in self.eulerian_fields.items():
                    stored = data[f"Eulerian/{self.eulerian_fields_type[name]}/{name}"]
                    field[inner] = np.ascontiguousarray(stored[self.local_eulerian_index])
            for grid_name, names in self.lagrangian_fields_with_grid_name.items():
                if mc.rank != self.lagrangian_grid_master_rank[grid_name]:
                    continue
                self.lagrangian_grids[grid_name][...] = data[f"Lagrangian/{grid_name}/Grid"]
                for name in names:
                    self.lagrangian_fields[name][...] = data[
                        f"Lagrangian/{grid_name}/{self.lagrangian_fields_type[name]}/{name}"]
        mc.grid.Barrier()
        return time

    # XDMF companions need the HDF5 container; kept as no-ops for call-site compatibility
    def generate_xdmf_eulerian(self, h5_file_name, time=0.0):
        pass

    def generate_xdmf_lagrangian(self, h5_file_name, time=0.0):
        pass
