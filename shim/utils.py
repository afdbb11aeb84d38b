"""Drop-in for the reference's `utils` module: every name train.py:20-21, eval.py:6 and configs/config.py:136 import
(`init_weights, Calculate_D_steps, plot_scores, plot_grad_norm, Checkpointer, ValidatedInput, sample_latent_vec,
plot_gen_samples`) plus the helpers around them."""
from neuron_gan_b200.utils import (Calculate_D_steps, Checkpointer, Latent_vecs_memo, N_params, ValidatedInput,  # noqa: F401
                                   calculate_grad_norm_hist, gen_samples, get_saved_attrs, init_weights,
                                   plot_gen_samples, plot_grad_norm, plot_scores, sample_latent_vec, save_vars,
                                   set_saved_attrs)
