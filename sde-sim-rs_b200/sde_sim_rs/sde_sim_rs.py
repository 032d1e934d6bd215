"""`sde_sim_rs.sde_sim_rs` — the module path of the reference's compiled extension (pyproject.toml:32 `module-name =
"sde_sim_rs.sde_sim_rs"`, re-exported by python/sde_sim_rs/__init__.py:1 `from .sde_sim_rs import simulate`), so that
`from sde_sim_rs.sde_sim_rs import simulate` keeps working.  Same function as `sde_sim_rs.simulate`."""
from . import simulate

__all__ = ["simulate"]
