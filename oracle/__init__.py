"""CPU oracle of the hot path -- TEST INFRASTRUCTURE, never imported by the product package."""
