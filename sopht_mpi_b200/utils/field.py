"""Vector-field component ordering, mirrors ``sopht.utils.field.VectorField``
as used by the reference (``sopht_mpi/utils/mpi_utils_3d.py:1125-1134``)."""


class VectorField:
    @staticmethod
    def x_axis_idx() -> int:
        return 0

    @staticmethod
    def y_axis_idx() -> int:
        return 1

    @staticmethod
    def z_axis_idx() -> int:
        return 2
