"""Import shim (test infrastructure only): minimal `dask` stub so the reference's
modules import; no dask graph is ever built through it."""
