"""Headless stand-in for the reference's vis/ba_vis.py.

``ba.py:79-81,103`` builds a scene and a viewer unconditionally.  The viewer (trimesh + pyglet)
is UI and out of scope; these no-op objects keep the unmodified script running.  ``update``
still touches ``cam.mu`` / ``lmk.mu`` like the real viewer (vis/ba_vis.py:35-55) so the
per-iteration device->host read of the means is exercised."""


class _Camera:
    resolution = (640, 480)


class _Scene:
    camera = _Camera()


def create_scene(graph, fov=(640, 480)):
    return _Scene()


class TrimeshSceneViewer:
    def __init__(self, scene=None, resolution=None, **kwargs):
        self.scene, self.resolution = scene, resolution
        self.n_updates = 0

    def show(self):
        return None

    def update(self, graph):
        if len(graph.cam_nodes):
            graph.cam_nodes[0].mu
        if len(graph.lmk_nodes):
            graph.lmk_nodes[0].mu
        self.n_updates += 1
