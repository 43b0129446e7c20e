"""GPI-PD's learned dynamics model (reference morl_baselines/common/model_based/)."""
