"""Model registry with the reference's keys (tapqir/models/__init__.py:17-21)."""

from tapqir_b200.models.cosmos import cosmos
from tapqir_b200.models.hmm import hmm
from tapqir_b200.models.model import Model

__all__ = ["models", "Model", "cosmos", "hmm"]

models = {cosmos.name: cosmos, hmm.name: hmm}
