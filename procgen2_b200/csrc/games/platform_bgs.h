// The 49 platformer background images shared by coinrun, jumper and climber
// (games/coinrun/coinrun.cpp:60-110; the other two lists are byte-identical).
#pragma once
#define PG2_PLATFORM_BACKGROUNDS \
    "assets/platform_backgrounds/alien_bg.png", "assets/platform_backgrounds/another_world_bg.png", \
    "assets/platform_backgrounds/back_cave.png", "assets/platform_backgrounds/caverns.png", \
    "assets/platform_backgrounds/cyberpunk_bg.png", "assets/platform_backgrounds/parallax_forest.png", \
    "assets/platform_backgrounds/scifi_bg.png", "assets/platform_backgrounds/scifi2_bg.png", \
    "assets/platform_backgrounds/living_tissue_bg.png", "assets/platform_backgrounds/airadventurelevel1.png", \
    "assets/platform_backgrounds/airadventurelevel2.png", "assets/platform_backgrounds/airadventurelevel3.png", \
    "assets/platform_backgrounds/airadventurelevel4.png", "assets/platform_backgrounds/cave_background.png", \
    "assets/platform_backgrounds/blue_desert.png", "assets/platform_backgrounds/blue_grass.png", \
    "assets/platform_backgrounds/blue_land.png", "assets/platform_backgrounds/blue_shroom.png", \
    "assets/platform_backgrounds/colored_desert.png", "assets/platform_backgrounds/colored_grass.png", \
    "assets/platform_backgrounds/colored_land.png", "assets/platform_backgrounds/colored_shroom.png", \
    "assets/platform_backgrounds/landscape1.png", "assets/platform_backgrounds/landscape2.png", \
    "assets/platform_backgrounds/landscape3.png", "assets/platform_backgrounds/landscape4.png", \
    "assets/platform_backgrounds/battleback1.png", "assets/platform_backgrounds/battleback2.png", \
    "assets/platform_backgrounds/battleback3.png", "assets/platform_backgrounds/battleback4.png", \
    "assets/platform_backgrounds/battleback5.png", "assets/platform_backgrounds/battleback6.png", \
    "assets/platform_backgrounds/battleback7.png", "assets/platform_backgrounds/battleback8.png", \
    "assets/platform_backgrounds/battleback9.png", "assets/platform_backgrounds/battleback10.png", \
    "assets/platform_backgrounds/sunrise.png", "assets/platform_backgrounds_2/beach1.png", \
    "assets/platform_backgrounds_2/beach2.png", "assets/platform_backgrounds_2/beach3.png", \
    "assets/platform_backgrounds_2/beach4.png", "assets/platform_backgrounds_2/fantasy1.png", \
    "assets/platform_backgrounds_2/fantasy2.png", "assets/platform_backgrounds_2/fantasy3.png", \
    "assets/platform_backgrounds_2/fantasy4.png", "assets/platform_backgrounds_2/candy1.png", \
    "assets/platform_backgrounds_2/candy2.png", "assets/platform_backgrounds_2/candy3.png", \
    "assets/platform_backgrounds_2/candy4.png"
#define PG2_NUM_PLATFORM_BACKGROUNDS 49
