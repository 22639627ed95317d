"""pinns_b200: B200-native (sm_100a) PINN training hot path behind the reference's
`PhysicsInformedNN` class API (jonwittmer/PINNs).  See DESIGN.md."""
from .engine import Engine  # noqa: F401

__all__ = ["Engine"]
