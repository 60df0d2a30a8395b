python tools/_slack.py
