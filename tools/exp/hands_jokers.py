"""Throughput of the general hand-scoring kernel (joker interpreter): 2^22 hands, 5 random jokers,
config-3 modifiers, n_cards 1..8 — next to the static 5-card kernel."""
import torch
from balatro_gym_b200.score import score_hands
dev = torch.device("cuda"); n = 1 << 22
g = torch.Generator(device=dev).manual_seed(1)
keys = torch.rand((n, 52), device=dev, generator=g)
cards = keys.topk(8, dim=1).indices.to(torch.uint8).contiguous()
ncards = torch.randint(1, 9, (n,), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
jk = (torch.rand((n, 145), device=dev, generator=g).topk(5, dim=1).indices + 1).to(torch.uint8)
jokers = torch.zeros((n, 8), dtype=torch.uint8, device=dev); jokers[:, :5] = jk
enh = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.25, torch.randint(1, 9, (n, 8), device=dev, generator=g), torch.zeros((n, 8), dtype=torch.int64, device=dev))
ed = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.1, torch.randint(1, 4, (n, 8), device=dev, generator=g), torch.zeros((n, 8), dtype=torch.int64, device=dev))
seal = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.1, torch.randint(1, 5, (n, 8), device=dev, generator=g), torch.zeros((n, 8), dtype=torch.int64, device=dev))
m = (enh << 6) | (ed << 10) | (seal << 13)
mods = torch.where(m >= 2 ** 15, m - 2 ** 16, m).to(torch.int16).contiguous()
levels = torch.randint(1, 6, (n, 12), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = score_hands(cards, mods8=mods, n_cards=ncards, jokers8=jokers, levels12=levels)
t = timed(lambda: score_hands(cards, mods8=mods, n_cards=ncards, jokers8=jokers, levels12=levels, out=out))
print("general kernel: %.1f us for 2^22 hands = %.3e hands/s; at 64+21 B/hand -> %.2f TB/s" % (t * 1e3, n / t * 1e3, n * (8 + 16 + 1 + 8 + 12 + 1 + 4 + 4 + 8 + 8 + 4) / t / 1e9))
out5 = score_hands(cards, want_x_mult=False, want_money=False)
t5 = timed(lambda: score_hands(cards, want_x_mult=False, want_money=False, out=out5))
print("static 5-card kernel: %.1f us = %.3e hands/s" % (t5 * 1e3, n / t5 * 1e3))
