"""CPU check of the synchronisation protocol of the persistent attention kernel (uvltrack_b200/csrc/attention3.cuh):
tools/attn3_protocol_sim.py replays the mbarrier phases of every role of one CTA with random completion delays."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_attention3_protocol_drains_without_hazards():
    import attn3_protocol_sim as sim

    for n, BH, G in ((40, 3, 2), (129, 3, 1), (361, 24, 5), (553, 24, 5), (553, 3, 148), (1193, 3, 2)):
        geo = sim.Geo(n, BH, 1)
        g = min(G, geo.total)
        for cta in {0, g - 1}:
            for seed in range(3):
                sim.Sim(n, BH, 1, g, cta, seed).run()
