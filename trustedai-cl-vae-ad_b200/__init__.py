"""kcvae-b200: B200-native (sm_100a) KurtosisCVAE training and anomaly scoring behind the
reference's Python surface (gtemplin/TrustedAI-CL-VAE-AD src/load_model.py, src/*_cvae.py)."""
from .load_model import (import_vae_based_on_type, load_config, load_model_from_config,  # noqa: F401
                         load_model_from_config_path, load_model_from_directory, save_config,
                         save_model_to_directory)
from .model import (AbstractCVAE, BetaAnnealingCallback, Callback, KTensor, KurtosisGlobalCVAE,  # noqa: F401
                    KurtosisSingleCVAE)
from .optimizers import Adam  # noqa: F401
from .scoring import evaluate_anomalies, get_data_scale, output_anomalies, rank_anomalies  # noqa: F401
from .streaming import DeviceDataQueue, StreamingAnomalyScore, render_outputs  # noqa: F401
