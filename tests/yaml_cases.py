"""YAML inputs shared by the oracle-parser tests and the native front-end tests.

ERROR_CASES follow the intent of the reference's own parser tests
(/root/reference/tests/test_sandengine-lang.rs:3-166 -- six error-path tests, stale against the
checked-in parser) and add the remaining error sites of the parser.
Each entry: (name, yaml, expected ParsingErr class).
"""

BASE_OK = """
rules:
  gravity:
    if: DOWN.mat.density < SELF.mat.density
    do: SWAP SELF DOWN
    mirrored: false
  slide_diagonally:
    if: DOWNRIGHT.mat.density < SELF.mat.density
    do: SWAP SELF DOWNRIGHT
    mirrored: true
types:
  movable_solid:
    base_rules: [gravity, slide_diagonally]
materials:
  sand:
    color: [1.0, 1.0, 0.0, 1.0]
    type: movable_solid
    density: 1.5
    selectable: true
"""

ERROR_CASES = [
    # test_sandengine-lang.rs:3-22 (missing_rules)
    ("missing_rules", """
types:
  movable_solid:
    base_rules: [gravity]
materials:
  sand: {color: [1.0, 1.0, 0.0, 1.0], type: movable_solid, density: 1.5}
""", "MissingField"),
    # :25-46 (missing_types)
    ("missing_types", """
rules:
  gravity: {if: DOWN.mat.density < SELF.mat.density, do: SWAP SELF DOWN}
materials:
  sand: {color: [1.0, 1.0, 0.0, 1.0], type: movable_solid, density: 1.5}
""", "MissingField"),
    # :49-71 (missing_materials)
    ("missing_materials", """
rules:
  gravity: {if: DOWN.mat.density < SELF.mat.density, do: SWAP SELF DOWN}
types:
  movable_solid:
    base_rules: [gravity]
""", "MissingField"),
    # :74-88 (invalid_name): a float key.  With all three sections present the InvalidType is reached.
    ("invalid_name", """
rules:
  1.0:
    if: DOWN.mat.density < SELF.mat.density
    do: SWAP SELF DOWN
types:
  solid:
materials:
  sand: {color: [1.0, 1.0, 0.0, 1.0], type: solid, density: 1.5}
""", "InvalidType"),
    # :90-103 (missing_field, first half): no `if`
    ("missing_if", """
rules:
  gravity:
    do: SWAP SELF DOWN
    mirrored: false
types:
  solid:
materials:
  sand: {color: [1.0, 1.0, 0.0, 1.0], type: solid, density: 1.5}
""", "MissingField"),
    # :105-132 (missing_field, second half): no `color`
    ("missing_color", BASE_OK.replace("    color: [1.0, 1.0, 0.0, 1.0]\n", ""), "MissingField"),
    # :136-166 (not_found): material type undefined
    ("type_not_found", BASE_OK.replace("type: movable_solid", "type: liquid"), "NotFound"),
    # function-call syntax from the README is not recognised (rules.rs:412-417)
    ("swap_call_syntax", BASE_OK.replace("do: SWAP SELF DOWNRIGHT", "do: swap(SELF, DOWNRIGHT)"), "NotRecognized"),
    ("missing_do", BASE_OK.replace("    do: SWAP SELF DOWN\n", ""), "MissingField"),
    ("do_bad_cell", BASE_OK.replace("do: SWAP SELF DOWN\n", "do: SWAP SELF UP\n"), "NotFound"),
    ("set_bad_cell", BASE_OK.replace("do: SWAP SELF DOWN\n", "do: SET UP sand\n"), "NotFound"),
    ("if_not_string", BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density", "if: 3"), "InvalidType"),
    ("mirrored_not_bool", BASE_OK.replace("mirrored: false", "mirrored: 1"), "InvalidType"),
    ("precondition_not_bool", BASE_OK.replace("mirrored: false", "precondition: maybe"), "InvalidType"),
    ("probability_not_float", BASE_OK.replace("mirrored: false", "probability: often"), "InvalidType"),
    ("unknown_material_in_if", BASE_OK.replace("if: DOWNRIGHT.mat.density < SELF.mat.density", "if: DOWN.mat == lava"), "NotFound"),
    ("unknown_type_in_if", BASE_OK.replace("if: DOWNRIGHT.mat.density < SELF.mat.density", "if: isType_lava(DOWN)"), "NotFound"),
    ("base_rule_not_found", BASE_OK.replace("base_rules: [gravity, slide_diagonally]", "base_rules: [gravity, nope]"), "NotFound"),
    ("base_rules_not_seq", BASE_OK.replace("base_rules: [gravity, slide_diagonally]", "base_rules: gravity"), "InvalidType"),
    ("inherits_not_string", BASE_OK.replace("    base_rules: [gravity, slide_diagonally]", "    inherits: [a]"), "InvalidType"),
    ("density_missing", BASE_OK.replace("    density: 1.5\n", ""), "MissingField"),
    ("density_not_number", BASE_OK.replace("density: 1.5", "density: heavy"), "InvalidType"),
    ("color_two_components", BASE_OK.replace("color: [1.0, 1.0, 0.0, 1.0]", "color: [1.0, 1.0]"), "InvalidType"),
    ("color_out_of_range", BASE_OK.replace("color: [1.0, 1.0, 0.0, 1.0]", "color: [300, 1.0, 0.0]"), "InvalidType"),
    ("color_not_seq", BASE_OK.replace("color: [1.0, 1.0, 0.0, 1.0]", "color: red"), "InvalidType"),
    ("extra_rules_not_seq", BASE_OK.replace("    selectable: true", "    extra_rules: gravity"), "InvalidType"),
    ("rules_not_mapping", "rules: [a, b]\ntypes: {}\nmaterials: {}\n", "InvalidType"),
    ("material_type_missing", BASE_OK.replace("    type: movable_solid\n", ""), "MissingField"),
    # inheritance cycles overflow the reference's stack (types.rs:186-210 recurse unguarded); a clean error here
    ("inherits_itself", BASE_OK.replace("    base_rules: [gravity, slide_diagonally]", "    inherits: movable_solid\n    base_rules: [gravity, slide_diagonally]"), "InvalidType"),
    ("inheritance_cycle", BASE_OK.replace("  movable_solid:\n    base_rules: [gravity, slide_diagonally]",
                                          "  aa:\n    inherits: bb\n  bb:\n    inherits: aa\n  movable_solid:\n    base_rules: [gravity, slide_diagonally]"), "InvalidType"),
    # LEFT handling defined by this build
    ("left_in_mirrored_rule", BASE_OK.replace("if: DOWNRIGHT.mat.density < SELF.mat.density\n    do: SWAP SELF DOWNRIGHT",
                                              "if: DOWNLEFT.mat.density < SELF.mat.density\n    do: SWAP SELF DOWNLEFT"), "NotRecognized"),
    ("left_mixed_with_right", BASE_OK.replace("if: DOWN.mat.density < SELF.mat.density\n    do: SWAP SELF DOWN\n    mirrored: false",
                                              "if: LEFT.mat.density < RIGHT.mat.density\n    do: SWAP SELF LEFT\n    mirrored: false"), "NotRecognized"),
]

# Accepted since round 2: a conflicting rule nobody uses does not reject the file (the reference emits used rules
# only).  (LEFT / RIGHT are matched as whole identifiers for this check; the reference's own literal replaces -- rules.rs:338-351 --
# still mangle any name that contains them, so such names stay unusable exactly as in the reference.)
UNUSED_LEFT_CONFLICT_OK = BASE_OK.replace("types:", """  draft_never_used:
    if: LEFT.mat.density < RIGHT.mat.density
    do: SWAP SELF LEFT
    mirrored: false
types:""")

# A rule set exercising most of the language: nested else chains, trailing else without `if`, lists of
# actions, SET+SWAP in one string, deep inheritance, non-mirrored RIGHT and LEFT rules, density vs
# literal, .mat.type / TYPE_ constants, != on materials, several probabilities.
RICH_YAML = """
rules:
  sink:
    if: DOWN.mat.density < SELF.mat.density and not isType_static(DOWN)
    do: SWAP SELF DOWN
    else:
      if: DOWNRIGHT.mat.density < SELF.mat.density and RIGHT.mat.density < SELF.mat.density
      probability: 0.5
      do: SWAP SELF DOWNRIGHT
      else:
        if: isType_fluid(SELF) and RIGHT.mat.density < SELF.mat.density
        probability: 0.75
        do: SWAP SELF RIGHT
  drift_right:
    mirrored: false
    if: isType_gasish(SELF) and isType_EMPTY(RIGHT)
    probability: 0.25
    do: SWAP SELF RIGHT
  drift_left:
    mirrored: false
    if: isType_gasish(SELF) and isType_EMPTY(LEFT)
    probability: 0.125
    do: SWAP SELF LEFT
  creep_left:
    mirrored: false
    precondition: false
    if: SELF.mat == moss and isType_EMPTY(DOWNLEFT) and LEFT.mat != m_rock
    probability: 0.05
    do: SET DOWNLEFT moss
  rise:
    precondition: false
    if: isType_gasish(DOWN) and DOWN.mat.density < SELF.mat.density and SELF.mat.type != TYPE_static
    do: SWAP DOWN SELF
    else:
      if: isType_gasish(DOWN) and DOWN.mat.density < RIGHT.mat.density and not isType_static(RIGHT)
      do: SWAP DOWN RIGHT
  burn:
    if: DOWN.mat == oil or RIGHT.mat == lava
    probability: 0.2
    do:
      - SET SELF fire
      - SWAP SELF DOWN
    else:
      do: SET SELF steam
      probability: 0.9
  cool:
    if: SELF.mat.density > 2.0 and RIGHT.mat == m_water
    do: SET SELF m_rock SWAP SELF RIGHT
    probability: 0.3
  condense:
    if: SELF.mat == steam and (rand.x < 0.25 or pos.y < 4)
    probability: 0.02
    do: SET SELF m_water
  spread:
    precondition: false
    if: isType_EMPTY(SELF) and DOWN.mat == moss and DOWNRIGHT.mat.density >= 1.0
    probability: 0.01
    do: SET SELF moss
types:
  static:
  granular:
    inherits: static
    base_rules: [sink]
  fine_granular:
    inherits: granular
  fluid:
    base_rules: [sink]
  thick_fluid:
    inherits: fluid
  gasish:
    base_rules: [rise, drift_right, drift_left]
  hot_gas:
    inherits: gasish
  plantish:
materials:
  m_rock: {type: static, color: [0.3, 0.3, 0.3], density: 6.0}
  gravel: {type: granular, color: [120, 120, 110], density: 2.2}
  dust: {type: fine_granular, color: [200, 190, 150], density: 1.6}
  m_water: {type: fluid, color: [0, 0, 1.0, 0.5], density: 1.2}
  oil: {type: thick_fluid, color: [0.2, 0.1, 0.0], density: 1.1}
  lava: {type: thick_fluid, color: [1.0, 0.3, 0.0], density: 2.8, emission: [1.0, 0.4, 0.1, 0.95], extra_rules: [cool]}
  steam: {type: gasish, color: [0.8, 0.8, 0.8, 0.4], density: 0.2, extra_rules: [condense]}
  fire: {type: hot_gas, color: [1.0, 0.6, 0.1], density: 0.05, emission: [1.0, 0.5, 0.0, 0.9], extra_rules: [burn]}
  moss: {type: plantish, color: [30, 120, 40], density: 2.0, selectable: false, extra_rules: [creep_left, spread]}
"""
RICH_IDS = {"EMPTY": 0, "m_rock": 3, "gravel": 4, "dust": 5, "m_water": 6, "oil": 7, "lava": 8, "steam": 9, "fire": 10, "moss": 11}
RICH_MIX = (("EMPTY", 0.40), ("gravel", 0.10), ("dust", 0.08), ("m_water", 0.12), ("oil", 0.06), ("lava", 0.04), ("steam", 0.06),
            ("fire", 0.04), ("moss", 0.05), ("m_rock", 0.05))

# Conditions that only "work by accident" in the reference (passed through verbatim as GLSL):
# arithmetic, integer modulo on pos, the `frame` uniform, rand.z / rand.w, TYPE_ constants, .mat.id, vector
# components of color / emission, nested parentheses, literal-on-the-left comparisons.
EXPR_YAML = """
rules:
  arith_fall:
    if: SELF.mat.density * 2 > DOWN.mat.density + 0.5 and (pos.x + 4) % 3 != 1
    do: SWAP SELF DOWN
    else:
      if: 1.25 < SELF.mat.density and (DOWNRIGHT.mat.density - SELF.mat.density) < -0.25 and RIGHT.mat.id < 3
      probability: 0.6
      do: SWAP SELF DOWNRIGHT
  frame_drift:
    mirrored: false
    if: frame % 2 == 0 and isType_EMPTY(RIGHT) and SELF.mat.type == TYPE_loose and (pos.y + 2) % 4 < 3
    do: SWAP SELF RIGHT
  glow_spread:
    precondition: false
    if: rand.z > 0.75 and isType_EMPTY(SELF) and DOWN.mat.emission.g > 0.5 and DOWN.mat.color.r <= 0.5
    do: SET SELF spark
  spark_fade:
    if: rand.w <= 0.5 or (rand.x >= 0.9 and frame % 3 == 0)
    do: SET SELF EMPTY
  id_swap:
    precondition: false
    if: SELF.mat.id >= 4 and DOWN.mat.id == 0 and not (RIGHT.mat != pebble) and pos.x >= 0
    probability: 0.3
    do: SWAP SELF DOWN
types:
  loose:
    base_rules: [arith_fall, frame_drift]
  glowing:
    base_rules: [glow_spread]
  shortlived:
    base_rules: [spark_fade]
materials:
  pebble: {type: loose, color: [0.5, 0.5, 0.5], density: 2.0, extra_rules: [id_swap]}
  ash:    {type: loose, color: [0.3, 0.3, 0.3], density: 1.25}
  ember:  {type: glowing, color: [0.4, 0.1, 0.1], density: 3.0, emission: [1.0, 0.6, 0.1, 0.9]}
  lamp:   {type: glowing, color: [0.9, 0.9, 0.2], density: 3.0, emission: [0.2, 0.9, 0.2, 0.8]}
  spark:  {type: shortlived, color: [1.0, 0.8, 0.2], density: 0.5, emission: [1.0, 0.8, 0.1, 0.7]}
"""
EXPR_IDS = {"EMPTY": 0, "pebble": 3, "ash": 4, "ember": 5, "lamp": 6, "spark": 7}
EXPR_MIX = (("EMPTY", 0.5), ("pebble", 0.15), ("ash", 0.15), ("ember", 0.08), ("lamp", 0.06), ("spark", 0.06))

# Scalar GLSL built-ins, integer bit operators, the conditional operator and vector == / != in conditions (the reference passes
# conditions to the GLSL compiler verbatim, so all of this is legal there).  Same materials as EXPR_YAML.
FUNC_YAML = """
rules:
  band_fall:
    if: abs(pos.x - 20) < 12 and DOWN.mat.density < SELF.mat.density
    do: SWAP SELF DOWN
    else:
      if: mod(float(pos.x + 1), 4.0) < 2.0 and min(RIGHT.mat.density, DOWNRIGHT.mat.density) < SELF.mat.density and max(pos.y, 3) > 3
      probability: 0.7
      do: SWAP SELF DOWNRIGHT
  parity_drift:
    mirrored: false
    if: ((pos.x + 1) & 1) == (frame & 1) and isType_EMPTY(RIGHT) and ((pos.y + 1) >> 1) % 3 != 0 and (frame ^ 5) > 1
    do: SWAP SELF RIGHT
  heat:
    precondition: false
    if: isType_EMPTY(SELF) and floor(DOWN.mat.density * 2.0) == 6.0 and clamp(rand.z, 0.25, 0.75) > 0.5 and fract(DOWN.mat.emission.r * 3.0) < 0.5
    do: SET SELF spark
  fade:
    if: "(rand.w <= 0.4 ? frame % 2 == 0 : step(0.8, rand.x) > 0.5) or sign(float(pos.x) - 30.5) * sqrt(float(frame % 16)) > 3.0"   # ': ' needs quotes in YAML
    do: SET SELF EMPTY
  glow_drop:
    precondition: false
    if: SELF.mat.emission.rgb != vec3(0.0) and DOWN.mat.color.rgba == vec4(0.0, 0.0, 0.0, 0.0) and RIGHT.mat.color.rg != vec2(0.5)
    probability: 0.8
    do: SWAP SELF DOWN
  sink_int:
    precondition: false
    if: int(SELF.mat.density * 2.0) >= 4 and DOWN.mat.id == 0 and ceil(SELF.mat.density) == 2.0 and (~pos.x | 1) != 0 and sign(pos.y - 5) >= 0
    probability: 0.5
    do: SWAP SELF DOWN
types:
  loose:
    base_rules: [band_fall, parity_drift]
  glowing:
    base_rules: [heat, glow_drop]
  shortlived:
    base_rules: [fade]
materials:
  pebble: {type: loose, color: [0.5, 0.5, 0.5], density: 2.0, extra_rules: [sink_int]}
  ash:    {type: loose, color: [0.3, 0.3, 0.3], density: 1.25}
  ember:  {type: glowing, color: [0.4, 0.1, 0.1], density: 3.0, emission: [1.0, 0.6, 0.1, 0.9]}
  lamp:   {type: glowing, color: [0.9, 0.9, 0.2], density: 3.0, emission: [0.2, 0.9, 0.2, 0.8]}
  spark:  {type: shortlived, color: [1.0, 0.8, 0.2], density: 0.5, emission: [1.0, 0.8, 0.1, 0.7]}
"""
FUNC_IDS, FUNC_MIX = EXPR_IDS, EXPR_MIX
