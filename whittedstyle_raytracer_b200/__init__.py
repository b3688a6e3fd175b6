"""whittedstyle_raytracer_b200 — B200-native Whitted ray-tracing core (drop-in for the
render hot path of bobhansky/WhittedStyle_Raytracer).  See DESIGN.md."""
from .scene import Scene, SceneError, read_ppm_p3, write_ppm_p3  # noqa: F401
from .renderer import Context, CudaError, CudaStrategy, MultiRenderer, Renderer  # noqa: F401
