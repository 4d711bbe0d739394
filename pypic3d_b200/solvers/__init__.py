from .first_order_yee import update_E, update_B  # noqa: F401
