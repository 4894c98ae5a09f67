"""placeholder - replaced below"""
def estimator_forward(*a, **k):
    raise RuntimeError("cvflow estimator binding not built yet")
