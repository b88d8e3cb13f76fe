from ...data.dataloader import DataLoader, build_data_loader, parallel_data_provider  # noqa: F401
