"""Tabular scalarised multi-objective Q-learning -- drop-in for reference morl_baselines/single_policy/ser/mo_q_learning.py
(``MOQLearning`` with the same constructor, ``eval / update / train / scalarized_q_values / get_config``).

BASELINE.json configs[0] ("MOQLearning on deep-sea-treasure-v0 (tabular, 2 objectives, CPU) -- plumbing, runs without a GPU"): this is
the one class of the package that needs no CUDA device.  It is not accelerated (component 28: a dictionary of
(|A|, d) float64 tables updated one transition at a time is host work); it exists so that the reference's CPU-runnable configuration
runs unchanged on the same API surface, and it reproduces the reference's tables bit for bit (tests/test_mo_q_learning_cpu.py against
tests/golden/moql.npz, frozen from the unmodified reference).

Dyna-Q (``dyna=True``: TabularModel, reference common/model_based/tabular_model.py) is outside the hot-path scope (SURVEY.md section 2,
component 20) and raises.
"""

from __future__ import annotations

import time
from collections.abc import Iterable
from typing import Optional

import numpy as np

from ...common.morl_algorithm import MOAgent, MOPolicy
from ...common.scalarization import weighted_sum
from ...common.utils import linearly_decaying_value


class MOQLearning(MOPolicy, MOAgent):
    """One Q-table per objective, actions chosen through a scalarisation function (Van Moffaert et al., ADPRL 2013)."""

    def __init__(self, env, id: Optional[int] = None, weights: np.ndarray = np.array([0.5, 0.5]), scalarization=weighted_sum,
                 learning_rate: float = 0.1, gamma: float = 0.9, initial_epsilon: float = 0.1, final_epsilon: float = 0.1,
                 epsilon_decay_steps: int = None, learning_starts: int = 0, use_gpi_policy: bool = False, dyna: bool = False,
                 dyna_updates: int = 5, model=None, gpi_pd: bool = False, min_priority: float = 0.0001, alpha: float = 0.6, parent=None,
                 project_name: str = "MORL-baselines", experiment_name: str = "MO Q-Learning", wandb_entity: Optional[str] = None,
                 log: bool = True, seed: Optional[int] = None, parent_rng: Optional[np.random.Generator] = None):
        """Same arguments, in the same order, as the reference constructor (mo_q_learning.py:26-52)."""
        MOAgent.__init__(self, env, device="cpu")
        MOPolicy.__init__(self, id, device="cpu")
        if dyna or model is not None:
            raise NotImplementedError("dyna=True (tabular Dyna-Q model) is outside the hot-path scope (SURVEY.md section 2, component 20)")
        self.id, self.seed, self.parent, self.log = id, seed, parent, log
        self.idstr = "" if id is None else f"_{id}"
        self.np_random = np.random.default_rng(seed) if parent_rng is None else parent_rng
        self.learning_rate, self.gamma = learning_rate, gamma
        self.initial_epsilon, self.final_epsilon, self.epsilon_decay_steps = initial_epsilon, final_epsilon, epsilon_decay_steps
        self.epsilon = initial_epsilon
        self.learning_starts, self.use_gpi_policy = learning_starts, use_gpi_policy
        self.dyna, self.dyna_updates, self.model = False, dyna_updates, None
        self.gpi_pd, self.min_priority, self.alpha = gpi_pd, min_priority, alpha
        self.weights, self.scalarization = weights, scalarization
        self.q_table = {}  # state tuple -> float64 [|A|, d]
        if log and parent_rng is None:
            self.setup_wandb(project_name, experiment_name, wandb_entity)

    def _act(self, obs) -> int:
        """Epsilon-greedy action (reference mo_q_learning.py:123-129): one draw of the policy's generator per step."""
        if self.np_random.random() < self.epsilon:
            return int(self.env.action_space.sample())
        return self.eval(obs, self.weights)

    @staticmethod
    def _state_to_tuple(obs) -> tuple:
        return tuple(obs) if isinstance(obs, Iterable) else (obs,)

    def scalarized_q_values(self, obs, w: np.ndarray) -> np.ndarray:
        """Scalarised Q value of every action for the observation and weights (reference :137-142)."""
        t_obs = self._state_to_tuple(obs)
        if t_obs not in self.q_table:
            return np.zeros(self.action_dim)
        return np.array([self.scalarization(v, w) for v in self.q_table[t_obs]])

    def eval(self, obs, w: Optional[np.ndarray] = None) -> int:
        """Greedy action under the policy's own weights (reference :160-171: ``w`` is only forwarded to a GPI parent); a state never
        visited gets a random action from the environment's sampler."""
        if self.use_gpi_policy:
            return self.parent.eval(obs, w)
        t_obs = self._state_to_tuple(obs)
        if t_obs not in self.q_table:
            return int(self.env.action_space.sample())
        return int(np.argmax(np.array([self.scalarization(v, self.weights) for v in self.q_table[t_obs]])))

    def update(self):
        """One tabular TD step on the transition stored in ``self.obs / action / reward / next_obs / terminated`` (reference :173-227)."""
        obs, next_obs = self._state_to_tuple(self.obs), self._state_to_tuple(self.next_obs)
        for s in (obs, next_obs):
            if s not in self.q_table:
                self.q_table[s] = np.zeros((self.action_dim, self.reward_dim))
        max_q = self.q_table[next_obs][self.eval(self.next_obs, self.weights)]
        td_error = self.reward + (1 - self.terminated) * self.gamma * max_q - self.q_table[obs][self.action]
        self.q_table[obs][self.action] += self.learning_rate * td_error
        if self.epsilon_decay_steps is not None:
            self.epsilon = linearly_decaying_value(self.initial_epsilon, self.epsilon_decay_steps, self.global_step, self.learning_starts,
                                                   self.final_epsilon)
        if self.log and self.global_step % 1000 == 0:
            import wandb

            wandb.log({f"charts{self.idstr}/epsilon": self.epsilon,
                       f"losses{self.idstr}/scalarized_td_error": self.scalarization(td_error, self.weights),
                       f"losses{self.idstr}/mean_td_error": np.mean(td_error), "global_step": self.global_step})

    def get_config(self) -> dict:
        return {"env_id": self.env.unwrapped.spec.id, "learning_rate": self.learning_rate, "gamma": self.gamma,
                "initial_epsilon": self.initial_epsilon, "final_epsilon": self.final_epsilon, "epsilon_decay_steps": self.epsilon_decay_steps,
                "use_gpi_policy": self.use_gpi_policy, "dyna": self.dyna, "dyna_updates": self.dyna_updates, "gpi_pd": self.gpi_pd,
                "min_priority": self.min_priority, "alpha": self.alpha, "weight": self.weights, "scalarization": self.scalarization.__name__,
                "seed": self.seed}

    def train(self, start_time, total_timesteps: int = int(5e5), reset_num_timesteps: bool = True, eval_env=None, eval_freq: int = 1000):
        """Interaction loop (reference :249-311): act, step, update, reset at episode ends."""
        self.obs, _ = self.env.reset()
        self.global_step = 0 if reset_num_timesteps else self.global_step
        self.num_episodes = 0 if reset_num_timesteps else self.num_episodes
        for _ in range(1, total_timesteps + 1):
            self.global_step += 1
            self.action = self._act(self.obs)
            self.next_obs, self.reward, self.terminated, self.truncated, info = self.env.step(self.action)
            self.update()
            if eval_env is not None and self.log and self.global_step % eval_freq == 0:
                self.policy_eval(eval_env, scalarization=self.scalarization, weights=self.weights, log=self.log)
            if self.terminated or self.truncated:
                self.obs, _ = self.env.reset()
                self.num_episodes += 1
                if self.log and self.global_step % 1000 == 0:
                    import wandb

                    wandb.log({f"charts{self.idstr}/SPS": int(self.global_step / (time.time() - start_time)), "global_step": self.global_step})
                    if "episode" in info:
                        from ...common.evaluation import log_episode_info

                        log_episode_info(info["episode"], self.scalarization, self.weights, self.global_step, self.id, verbose=False)
            else:
                self.obs = self.next_obs
