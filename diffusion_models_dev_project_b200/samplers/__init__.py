from .base_sampler import BaseSampler
from .adaptation import tv_loss, adaptation_loss, AdaptationLoss, adapt_objective, adapt_objective_applies, _score_model_adpt
from .utils import (apTweedy, ddim, decomposed_diffusion_sampling_sde_predictor,
                    adapted_ddim_sde_predictor, _adapt, _schedule_jump, wrapper_ddim,
                    Euler_Maruyama_sde_predictor, Ancestral_Sampling, Langevin_sde_corrector)
