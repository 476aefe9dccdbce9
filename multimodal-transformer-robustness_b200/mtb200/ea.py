"""Evolutionary search over fusion sub-networks (reference: EA.py:21-169), B200 edition.

Same genome ([active_cross, active_cross_output]), same mutate / crossover / parent selection
and the same consumption of the three host RNG streams (torch CPU generator, ``random``,
``numpy.random``) as the reference, so a seeded search visits the same candidates.  What
changes is WHERE fitness is computed:

* candidates of one generation are mutually independent (children only depend on the parents'
  stored scores), so each rank of a one-process-per-GPU job evaluates candidates r, r+N, ... and
  only the fp32 scores are all-gathered;
* in eval mode the per-modality `mems0` stacks and every cross-modal branch are identical for all
  candidates, so their outputs are memoised per validation batch (DynamicMULTModel.forward's
  ``branch_cache``) and a candidate only costs its masked `mems` stacks plus the head;
* the reference draws one int64 from the global CPU generator per evaluation (DataLoader
  iterator creation, EA.py:157); that draw is replayed so the sampling stream stays bit-exact.
"""
from __future__ import annotations

import copy
import random
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import dist as mdist

__all__ = ["EvolutionSearch", "binary_acc"]


def binary_acc(results: torch.Tensor, truths: torch.Tensor, exclude_zero: bool = True) -> float:
    """src/eval_metrics.py:18-25 (accuracy of sign agreement, optionally ignoring zero labels)."""
    p = results.reshape(-1)
    t = truths.reshape(-1)
    keep = (t != 0) if exclude_zero else torch.ones_like(t, dtype=torch.bool)
    if int(keep.sum()) == 0:
        return 0.0
    return float(((p[keep] > 0) == (t[keep] > 0)).float().mean())


class EvolutionSearch:
    def __init__(self, parent_model, hyper_params, valid_batches: Sequence, test_batches: Optional[Sequence] = None,
                 metric: Callable = binary_acc, memoize: bool = True):
        """valid_batches / test_batches: lists of ([x_modality...], y) already on the model's device."""
        hp = hyper_params
        self.mutate_prob = hp.mutate_prob
        self.population_size = hp.population_size
        self.max_time_budget = hp.max_time_budget
        self.parent_ratio = hp.parent_ratio
        self.mutation_ratio = hp.mutation_ratio
        self.active_modality = hp.active_modality
        self.hyper_params = hp
        self.valid_batches = list(valid_batches)
        self.test_batches = list(test_batches) if test_batches is not None else self.valid_batches
        self.model = parent_model
        self.metric = metric
        self.memoize = memoize
        self.latency_constraint = 100
        self._caches = {}
        self.evaluations = 0

    # ------------------------------------------------------------------ genome operators (EA.py:44-73)
    def mutate(self, sample):
        new_sample = copy.deepcopy(sample)
        probs = torch.rand(len(sample[1]),)
        for i in range(len(probs)):
            if probs[i] < self.mutate_prob:
                sub = self.model.gen_active_cross(active_modality=self.active_modality)
                new_sample[0][i] = copy.deepcopy(sub[0][i])
                new_sample[1][i] = copy.deepcopy(sub[1][i])
        return new_sample, 0

    def crossover(self, sample1, sample2):
        new_sample = copy.deepcopy(sample1)
        for i in range(len(new_sample[0])):
            if random.choice([0, 1]) == 0:
                new_sample[0][i] = copy.deepcopy(sample2[0][i])
                new_sample[1][i] = copy.deepcopy(sample2[1][i])
        return new_sample, 0

    # ------------------------------------------------------------------ fitness
    def reset_memo(self):
        """drop the memoised branch outputs (new weights or new validation data)"""
        self._caches.clear()

    def _replay_loader_draw(self):
        torch.empty((), dtype=torch.int64).random_()      # the DataLoader-iterator draw of EA.py:157

    def predict(self, test: bool = False, only_batch: Optional[int] = None):
        """One inference pass over the (validation | test) batches (or just batch `only_batch`) with the model's current
        config; returns the device predictions per batch WITHOUT reading them back (no host synchronisation)."""
        self.model.eval()
        batches = self.test_batches if test else self.valid_batches
        outs = []
        with torch.no_grad():
            for bi, (xs, y) in enumerate(batches):
                if only_batch is not None and bi != only_batch:
                    continue
                cache = self._caches.setdefault((test, bi), {}) if self.memoize else None
                pred, _ = self.model(xs, branch_cache=cache) if cache is not None else self.model(xs)
                outs.append(pred)
        if only_batch is None or only_batch == len(batches) - 1:
            self.evaluations += 1
        return outs

    def _truths(self, test: bool = False) -> torch.Tensor:
        key = ("truths", test)
        t = self.__dict__.get("_truth_cache", {}).get(key)
        if t is None:
            batches = self.test_batches if test else self.valid_batches
            t = torch.cat([y for _, y in batches]).float().cpu()
            self.__dict__.setdefault("_truth_cache", {})[key] = t
        return t

    def eval_model(self, test: bool = False) -> float:
        """EA.py:149-169: predictions over the whole loader, then the metric on the host."""
        outs = self.predict(test)
        return self.metric(torch.cat(outs).float().cpu(), self._truths(test))

    def get_acc(self, sample) -> float:
        self.model.set_active_modalities(active_modality=self.active_modality, active_cross=copy.deepcopy(sample[0]),
                                         active_cross_output=copy.deepcopy(sample[1]))
        return self.eval_model()

    def score_many(self, samples: List) -> List[float]:
        """Fitness of a list of candidates, sharded across ranks (candidate r, r + N, ...); only the scores are all-reduced.
        The forwards of ALL of this rank's candidates are enqueued before the first prediction is read back, so building
        the next candidate's plan on the host overlaps the GPU work of the previous ones (the reference's get_acc reads
        every score back before it configures the next candidate, EA.py:75-81)."""
        dev = next(self.model.parameters()).device
        rank, n = mdist.world()
        mine = mdist.shard_indices(len(samples), rank, n)
        pending = {i: [] for i in mine}
        # batch-major: the memoised branch outputs live in the engine's regions for ONE validation batch at a time, so all
        # candidates are run on batch 0, then all on batch 1, ...
        for bi in range(len(self.valid_batches)):
            for i in mine:
                self.model.set_active_modalities(active_modality=self.active_modality, active_cross=copy.deepcopy(samples[i][0]),
                                                 active_cross_output=copy.deepcopy(samples[i][1]))
                pending[i].extend(self.predict(only_batch=bi))
        scores = torch.zeros(len(samples), dtype=torch.float32, device=dev)
        truths = self._truths()
        for i in mine:
            scores[i] = float(self.metric(torch.cat(pending[i]).float().cpu(), truths))
        if n > 1:
            import torch.distributed as dist
            dist.all_reduce(scores, op=dist.ReduceOp.SUM)
        return scores.tolist()

    # ------------------------------------------------------------------ search (EA.py:84-137)
    def search(self):
        mutation_numbers = int(round(self.mutation_ratio * self.population_size))
        parents_size = int(round(self.parent_ratio * self.population_size))
        best_valids = [-10]
        best_info = None
        samples = []
        for _ in range(self.population_size):
            cross, outs = self.model.gen_active_cross(active_modality=self.active_modality)
            samples.append([cross, outs])
            self._replay_loader_draw()
        population = [[a, s] for a, s in zip(self.score_many(samples), samples)]
        for it in range(self.max_time_budget):
            parents = sorted(population, key=lambda x: x[0])[::-1][:parents_size]
            acc = parents[0][0]
            if acc > best_valids[-1]:
                best_valids.append(acc)
                best_info = copy.deepcopy(parents[0])
            else:
                best_valids.append(best_valids[-1])
            if it >= self.max_time_budget - 1:
                self.model.set_active_modalities(active_modality=self.active_modality, active_cross=best_info[1][0],
                                                 active_cross_output=best_info[1][1])
                return best_valids, best_info
            population = copy.deepcopy(parents)
            children = []
            for _ in range(mutation_numbers):
                par = population[np.random.randint(parents_size)][1]
                child, _ = self.mutate(par)
                children.append(child)
                self._replay_loader_draw()
            for _ in range(self.population_size - mutation_numbers):
                p1 = population[np.random.randint(parents_size)][1]
                p2 = population[np.random.randint(parents_size)][1]
                child, _ = self.crossover(p1, p2)
                children.append(child)
                self._replay_loader_draw()
            population += [[a, s] for a, s in zip(self.score_many(children), children)]
        return best_valids, best_info

    def test_modality(self, active_code):
        self.model.set_active_modalities(active_modality=self.active_modality, active_cross=active_code[0],
                                         active_cross_output=active_code[1])
        self.eval_model()
        return self.eval_model(test=True)
