"""Create an evaluation, push samples in size-adaptive batches, finalize with metrics. Needs PRIME_API_KEY.

The sample dicts follow the verifiers results format (one row per rollout); unknown keys are preserved by the hub.
"""

import random

from prime_b200.platform.evals import APIClient, EvalsClient

client = EvalsClient(APIClient())
ev = client.create_evaluation(
    name="gsm8k-demo-run",
    environments=["primeintellect/gsm8k"],  # slug → looked up; a bare name would be resolved get-or-create
    model_name="meta-llama/Llama-3.1-8B-Instruct",
    framework="verifiers",
    task_type="math",
    metadata={"num_examples": 32, "rollouts_per_example": 2, "sampling_args": {"temperature": 0.7, "max_tokens": 512}},
)
eval_id = ev["evaluation_id"]
print("created", eval_id)

rng = random.Random(0)
samples = []
for ex in range(32):
    for rollout in range(2):
        ok = rng.random() < 0.6
        samples.append({
            "example_id": ex, "rollout_number": rollout, "task": "gsm8k", "reward": float(ok), "correct_answer": float(ok),
            "prompt": [{"role": "user", "content": f"Question #{ex}: what is {ex} + {ex}?"}],
            "completion": [{"role": "assistant", "content": f"The answer is {2 * ex if ok else 2 * ex + 1}."}],
            "answer": str(2 * ex), "info": {"tokens": rng.randint(20, 200)},
        })  # fmt: skip
pushed = client.push_samples(eval_id, samples)
print("pushed", pushed)
acc = sum(s["reward"] for s in samples) / len(samples)
print(client.finalize_evaluation(eval_id, metrics={"accuracy": acc, "avg_reward": acc, "num_samples": len(samples)}))
print("first page:", len(client.get_samples(eval_id, page=1, limit=10).get("samples", [])), "samples")
