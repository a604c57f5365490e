def convert_image_to_rgb(image, format):
    return image
