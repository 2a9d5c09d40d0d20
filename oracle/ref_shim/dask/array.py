"""Import shim (test infrastructure only): `dask.array.Array` type stub."""


class Array:  # noqa: D401 - stub
    pass
