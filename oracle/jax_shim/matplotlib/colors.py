hsv_to_rgb = rgb_to_hsv = ListedColormap = None
