"""TEST INFRASTRUCTURE: CPU oracle for the mvpraymarch path (never imported by the product package)."""
