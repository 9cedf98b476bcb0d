"""Headless stand-in for the reference's vis/vis_scene.py (trimesh/pyglet viewer: out of scope)."""


def view(*args, **kwargs):
    return None
