"""protnote_b200: the B200-native (sm_100a) ProtNote scoring path behind the reference's nn.Module interface."""
from ._lib import PN_FAST, PN_STRICT, ProtnoteB200Error  # noqa: F401
