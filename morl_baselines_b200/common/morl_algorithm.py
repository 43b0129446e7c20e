"""MORL base classes (mirrors reference morl_baselines/common/morl_algorithm.py -- the API boundary that is preserved).

``MOPolicy`` / ``MOAgent`` keep the reference's method names, arguments and attribute names.  Environment introspection
is duck-typed (``.n`` for discrete spaces, ``.shape`` otherwise) so it works with gymnasium spaces when gymnasium is
installed and with any look-alike otherwise; wandb is imported lazily and only when ``log=True``.
"""

from __future__ import annotations

import os
import time
from abc import ABC, abstractmethod
from typing import Dict, Optional, Union

import numpy as np
import torch as th

from .evaluation import eval_mo_reward_conditioned, policy_evaluation_mo


def _is_discrete(space) -> bool:
    return hasattr(space, "n") and not hasattr(space, "low")


class MOPolicy(ABC):
    """A multi-objective policy: greedy action via ``eval`` and learning via ``update`` (reference morl_algorithm.py:23-221)."""

    def __init__(self, id: Optional[int] = None, device: Union[th.device, str] = "auto") -> None:
        self.id = id
        self.device = th.device("cuda" if th.cuda.is_available() else "cpu") if device == "auto" else th.device(device)
        self.global_step = 0

    @abstractmethod
    def eval(self, obs: np.ndarray, w: Optional[np.ndarray]) -> Union[int, np.ndarray]:
        """Best action for the observation (and weight vector)."""

    def _report(self, scalarized_return, scalarized_discounted_return, vec_return, discounted_vec_return):
        import wandb

        idstr = "" if self.id is None else f"_{self.id}"
        wandb.log({f"eval{idstr}/scalarized_return": scalarized_return,
                   f"eval{idstr}/scalarized_discounted_return": scalarized_discounted_return,
                   "global_step": self.global_step})
        for i in range(vec_return.shape[0]):
            wandb.log({f"eval{idstr}/vec_{i}": vec_return[i], f"eval{idstr}/discounted_vec_{i}": discounted_vec_return[i]})

    def policy_eval(self, eval_env, num_episodes: int = 5, scalarization=np.dot, weights: Optional[np.ndarray] = None, log: bool = False):
        """Average returns over ``num_episodes`` evaluation episodes (reference morl_algorithm.py:85-126)."""
        res = policy_evaluation_mo(self, eval_env, scalarization=scalarization, w=weights, rep=num_episodes)
        if log:
            self._report(*res)
        return res

    def policy_eval_esr(self, eval_env, scalarization, weights: Optional[np.ndarray] = None, log: bool = False):
        """ESR evaluation on one episode (reference morl_algorithm.py:128-166)."""
        res = eval_mo_reward_conditioned(self, eval_env, scalarization, weights)
        if log:
            self._report(*res)
        return res

    def get_policy_net(self) -> th.nn.Module:
        pass

    def get_buffer(self):
        pass

    def set_buffer(self, buffer):
        pass

    def get_save_dict(self, save_replay_buffer: bool = False) -> dict:
        pass

    def save(self, save_dir: str = "weights/", filename: Optional[str] = None, save_replay_buffer: bool = False):
        os.makedirs(save_dir, exist_ok=True)
        filename = filename or f"policy_{self.id}.pth"
        th.save(self.get_save_dict(save_replay_buffer), os.path.join(save_dir, filename))

    def load(self, path, load_replay_buffer=True):
        pass

    def set_weights(self, weights: np.ndarray):
        pass

    @abstractmethod
    def update(self) -> None:
        """Update the policy's parameters."""


class MOAgent(ABC):
    """An agent holding one or more MOPolicies; extracts env features, seeds, sets up logging (reference :224-337)."""

    def __init__(self, env, device: Union[th.device, str] = "auto", seed: Optional[int] = None) -> None:
        self.extract_env_info(env)
        self.device = th.device("cuda" if th.cuda.is_available() else "cpu") if device == "auto" else th.device(device)
        self.global_step = 0
        self.num_episodes = 0
        self.seed = seed
        self.np_random = np.random.default_rng(self.seed)

    def extract_env_info(self, env) -> None:
        """Observation / action / reward dimensions of the environment (reference morl_algorithm.py:248-273)."""
        if env is None:
            return
        self.env = env
        if _is_discrete(env.observation_space):
            self.observation_shape = (1,)
            self.observation_dim = env.observation_space.n
        else:
            self.observation_shape = tuple(env.observation_space.shape)
            self.observation_dim = env.observation_space.shape[0]
        self.action_space = env.action_space
        if _is_discrete(env.action_space):
            self.action_shape = (1,)
            self.action_dim = env.action_space.n
        else:
            self.action_shape = tuple(env.action_space.shape)
            self.action_dim = env.action_space.shape[0]
        self.reward_dim = env.unwrapped.reward_space.shape[0]

    @abstractmethod
    def get_config(self) -> dict:
        """Algorithm hyper-parameters as a dictionary."""

    def register_additional_config(self, conf: Dict = {}) -> None:
        import wandb

        for key, value in conf.items():
            wandb.config[key] = value

    def setup_wandb(self, project_name: str, experiment_name: str, entity: Optional[str] = None, group: Optional[str] = None,
                    mode: Optional[str] = "online") -> None:
        """Initialise Weights & Biases with the reference's run naming and step metric (reference :292-331)."""
        import wandb

        self.experiment_name = experiment_name
        env0 = self.env.envs[0] if hasattr(self.env, "envs") else self.env
        self.full_experiment_name = f"{env0.spec.id}__{experiment_name}__{self.seed}__{int(time.time())}"
        config = self.get_config()
        config["algo"] = self.experiment_name
        monitor_gym = os.environ.get("MONITOR_GYM", "True").lower() in ("y", "yes", "t", "true", "on", "1")
        wandb.init(project=project_name, entity=entity, config=config, name=self.full_experiment_name, monitor_gym=monitor_gym,
                   save_code=True, group=group, mode=mode)
        wandb.define_metric("*", step_metric="global_step")

    def close_wandb(self) -> None:
        import wandb

        wandb.finish()
