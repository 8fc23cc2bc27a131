from .sde import SDE, VESDE, VPSDE, DDPM, _EPSILON_PRED_CLASSES, _SCORE_PRED_CLASSES
from .metrics import PSNR, SSIM
from .cg import cg
