"""`imp` was removed in Python 3.12; src/models/robot_model.py:1 imports it without using it."""
