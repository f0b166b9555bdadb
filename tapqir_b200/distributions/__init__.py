from tapqir_b200.distributions.affine_beta import AffineBeta
from tapqir_b200.distributions.ksmogn import KSMOGN

__all__ = ["AffineBeta", "KSMOGN"]
