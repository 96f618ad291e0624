from soccerdiffusion_b200.schedulers.scheduling_ddim import DDIMScheduler, DDIMSchedulerOutput  # noqa: F401
