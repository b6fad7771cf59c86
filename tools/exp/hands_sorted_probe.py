"""How much of the joker interpreter's time is divergence over the card count?  Same 2^22 hands, random order vs sorted by
n_cards (an upper bound for what binning tiles by card count could bring)."""
import torch
from balatro_gym_b200.score import score_hands
dev = torch.device("cuda:0")
n = 1 << 22
g = torch.Generator(device=dev); g.manual_seed(0xC3)
cards = torch.rand((n, 52), device=dev, generator=g).topk(8, dim=1).indices.to(torch.uint8).contiguous()
ncards = torch.randint(1, 9, (n,), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
jokers = torch.zeros((n, 8), dtype=torch.uint8, device=dev)
jokers[:, :5] = (torch.rand((n, 145), device=dev, generator=g).topk(5, dim=1).indices + 1).to(torch.uint8)
z = torch.zeros((n, 8), dtype=torch.int64, device=dev)
enh = torch.where(torch.rand((n, 8), device=dev, generator=g) < 0.25, torch.randint(1, 9, (n, 8), device=dev, generator=g), z)
m = (enh << 6)
mods = m.to(torch.int16).contiguous()
levels = torch.randint(1, 6, (n, 12), device=dev, generator=g, dtype=torch.int32).to(torch.uint8)
def timed(c, md, nc, jk, lv):
    out = score_hands(c, mods8=md, n_cards=nc, jokers8=jk, levels12=lv)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): score_hands(c, mods8=md, n_cards=nc, jokers8=jk, levels12=lv, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10
print("random order           %.3f ms" % timed(cards, mods, ncards, jokers, levels))
p = torch.argsort(ncards.int())
print("sorted by n_cards      %.3f ms" % timed(cards[p].contiguous(), mods[p].contiguous(), ncards[p].contiguous(), jokers[p].contiguous(), levels[p].contiguous()))
nc5 = torch.full_like(ncards, 5)
print("all 5 cards            %.3f ms" % timed(cards, mods, nc5, jokers, levels))
nc8 = torch.full_like(ncards, 8)
print("all 8 cards            %.3f ms" % timed(cards, mods, nc8, jokers, levels))
j0 = jokers.clone(); j0[:] = jokers[0]
print("same 5 jokers everywhere, random n_cards  %.3f ms" % timed(cards, mods, ncards, j0, levels))
