"""Deep Retrieval M-step on top of the C ABI (first half).

deep-retrieval/src/main/scala/com/mass/dr/optim/CoordinateDescent.scala:
  batchPathScore / streamingPathScore (:135-205)  every training sample's beam search over the K^D paths -- the work, done by
                                                  dmg_dr_beam_search for whole batches -- aggregated per target item
  optimize (:29-78)                               greedy coordinate descent: J paths per item maximising
                                                  n_v (log1p(p + partial) - log1p(partial)) - penalty(|path|)
The aggregation and the greedy loop are hash-map bookkeeping on the host, as in the reference.  Where the Scala code
iterates a HashMap (groupMapReduce(...).toSeq before a stable sort, idItemMapping.keys) the order of EQUAL scores is a JVM
artefact; here ties keep first-seen order and items are visited in ascending id -- the same choice for distinct scores.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

Path = Tuple[int, ...]
PathScores = List[Tuple[Path, float]]


def penalty_func(path_size: int, poly_order: int) -> float:
    """CoordinateDescent.penaltyFunc (:114-117): f(s+1) - f(s), f(s) = s^order / order."""
    f = lambda s: math.pow(s, poly_order) / poly_order
    return f(path_size + 1) - f(path_size)


def aggregate_path_score(num: int, per_sample: Sequence[PathScores]) -> PathScores:
    """aggregatePathScore (:119-126): sum of the probabilities per path over the item's samples (in sample order),
    stable sort by score descending, first `num`."""
    acc: Dict[Path, float] = {}
    for ps in per_sample:
        for path, prob in ps:
            acc[path] = acc[path] + prob if path in acc else prob
    items = list(acc.items())
    order = sorted(range(len(items)), key=lambda i: -items[i][1])          # Python's sort is stable
    return [items[i] for i in order[:num]]


def beam_search_batches(beam_search: Callable[[np.ndarray, int], Tuple[np.ndarray, np.ndarray, np.ndarray]], seqs: np.ndarray,
                        num_candidate_path: int, batch_size: int) -> List[PathScores]:
    """CandidateSearcher.beamSearch for every training sample, `batch_size` samples per engine call."""
    out: List[PathScores] = []
    for b0 in range(0, len(seqs), batch_size):
        paths, probs, counts = beam_search(seqs[b0:b0 + batch_size], num_candidate_path)
        for u in range(len(counts)):
            out.append([(tuple(int(x) for x in paths[u, q]), float(probs[u, q])) for q in range(int(counts[u]))])
    return out


def batch_path_score(beam_search, seqs: np.ndarray, targets: Sequence[int], num_candidate_path: int,
                     batch_size: int = 1024) -> Dict[int, PathScores]:
    """batchPathScore (:135-160)."""
    per_sample = beam_search_batches(beam_search, seqs, num_candidate_path, batch_size)
    by_item: Dict[int, List[PathScores]] = {}
    for t, ps in zip(targets, per_sample):
        by_item.setdefault(int(t), []).append(ps)
    return {item: aggregate_path_score(num_candidate_path, lst) for item, lst in by_item.items()}


def streaming_path_score(beam_search, seqs: np.ndarray, targets: Sequence[int], num_candidate_path: int, decay_factor: float,
                         batch_size: int = 1024) -> Dict[int, PathScores]:
    """streamingPathScore (:162-205): exponentially decayed path scores, samples in data order."""
    per_sample = beam_search_batches(beam_search, seqs, num_candidate_path, batch_size)
    scores: Dict[int, PathScores] = {}
    for t, cand in zip(targets, per_sample):
        item = int(t)
        if item not in scores:
            scores[item] = cand
            continue
        orig = scores[item]
        min_score = min(p for _, p in orig)
        o, c = dict(orig), dict(cand)
        union = list(o.keys()) + [p for p in c if p not in o]
        new = []
        for p in union:
            if p in o and p in c:
                s = decay_factor * o[p] + c[p]
            elif p in c:
                s = decay_factor * min_score + c[p]
            else:
                s = decay_factor * o[p]
            new.append((p, s))
        order = sorted(range(len(new)), key=lambda i: -new[i][1])
        scores[item] = [new[i] for i in order[:num_candidate_path]]
    return scores


def optimize(item_path_score: Dict[int, PathScores], item_occurrence: Dict[int, int], all_items: Sequence[int], num_iteration: int,
             num_path_per_item: int, num_layer: int, num_node: int, penalty_factor: float, penalty_poly_order: int,
             seed: int = 0) -> Dict[int, List[Path]]:
    """CoordinateDescent.optimize (:29-78).  Items without a training occurrence get random paths (generateRandomPath)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    mapping: Dict[int, List[Path]] = {}
    path_size: Dict[Path, int] = {}
    for t in range(1, num_iteration + 1):
        for v in sorted(int(x) for x in all_items):
            if v not in item_occurrence:
                mapping[v] = [tuple(int(x) for x in rng.integers(0, num_node, num_layer)) for _ in range(num_path_per_item)]
                continue
            selected: List[Path] = []
            partial = 0.0
            for j in reversed(range(num_path_per_item)):                 # foldRight over 0 until J: j = J-1 first
                if t > 1:
                    last = mapping[v][j]
                    path_size[last] = path_size.get(last, 0) - 1
                cands = [(p, pr) for p, pr in item_path_score[v] if p not in selected]
                if not cands:                                            # the reference's maxBy throws on an empty list too
                    raise ValueError(f"item {v}: fewer than {num_path_per_item} candidate paths")
                best, best_gain = None, -math.inf
                for p, pr in cands:                                      # maxBy: the first maximum wins
                    pen = penalty_factor * penalty_func(path_size.get(p, 0), penalty_poly_order)
                    gain = item_occurrence[v] * (math.log1p(pr + partial) - math.log1p(partial)) - pen
                    if gain > best_gain:
                        best, best_gain = p, gain
                path_size[best] = path_size.get(best, 0) + 1
                selected.insert(0, best)                                 # maxPath :: selectedPath
                partial += best_gain
            mapping[v] = selected
    return mapping
